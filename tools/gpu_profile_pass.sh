#!/bin/bash
# One profiling pass on the GPU box (B200_PROFILING.md recipe). Each ncu command runs only after the same command
# has exited 0 without ncu. Outputs land in gpurun_out/ (tag = $1).
tag=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -5 gpurun_out/bench_$tag.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_short.log 2>&1 || exit 1
# launch list (durations are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
# full captures of the two kernels that carry a roofline (one launch each, from the timed workload)
ncu --set full --import-source on --clock-control none -k regex:k_count_queue -s 4 -c 1 -o gpurun_out/prof_score_$tag -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_score_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_knn2_tc4 -s 4 -c 1 -o gpurun_out/prof_knn2tc_$tag -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_knn2tc_$tag.log 2>&1
cut -c1-400 gpurun_out/bench_$tag.log
python tools/ncu_traffic.py gpurun_out/prof_knn2tc_$tag.ncu-rep 1024 5000 1024 > gpurun_out/ncu_traffic_$tag.log 2>&1
