#!/bin/bash
# development aid: validate + time (+ optionally profile) the tensor-core Hamming kernel on the GPU box
mkdir -p gpurun_out
python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check exit $?" >> gpurun_out/tc_check.log
python tools/tc_time.py > gpurun_out/tc_time.log 2>&1
# (timing floors need a TUNING=1 build: option tc_dbg)
for m in 2 4 6; do VB_OPTIONS="tc_dbg=$m" python tools/tc_time.py 2>&1 | tail -1 > gpurun_out/tc_time_dbg$m.log; done
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_knn2_tc$ -c 1 -o gpurun_out/prof_tc6 -f python tools/tc_time.py > gpurun_out/ncu_tc6.log 2>&1
fi
cat gpurun_out/tc_check.log; tail -2 gpurun_out/tc_time.log; tail -n1 gpurun_out/tc_time_dbg*.log
