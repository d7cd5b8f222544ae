#!/bin/bash
# One profiling pass on the GPU box (B200_PROFILING.md recipe). Each ncu command runs only after the same command has
# exited 0 without ncu. Outputs land in gpurun_out/ (tag = $1); tools/ncu_summary.py turns the reports into the text
# summaries committed under profiles/.
tag=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_${tag}_short.log 2>&1 || { echo "bench failed"; tail -5 gpurun_out/bench_${tag}_short.log; exit 1; }
# launch list (durations are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
cap() {  # cap <kernel regex> <skip> <out name> <command...>
  local k=$1 s=$2 o=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s $s -c 1 -o gpurun_out/prof_${o}_$tag -f "$@" > gpurun_out/ncu_${o}_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${o}_$tag.ncu-rep > gpurun_out/${tag}_ncu_$o.txt 2>&1
}
cap k_knn2_tc4 4 k_knn2_tc4 $B
cap k_count_queue 4 k_count_queue $B
cap 'k_solve8$' 4 k_solve8 $B
cap k_knn2_tc_fix 4 k_knn2_tc_fix $B
cap k_expand_e2m1 4 k_expand_e2m1 $B
python tools/kd_profile.py > /dev/null 2>&1 && {
  cap 'k_kd_nearest' 3 k_kd_nearest python tools/kd_profile.py
  cap 'k_kd_radius' 6 k_kd_radius python tools/kd_profile.py
  cap 'k_kd_build' 2 k_kd_build python tools/kd_profile.py
  cap 'k_kd_split' 1 k_kd_split python tools/kd_profile.py
}
python tools/l2_time.py > /dev/null 2>&1 && {
  cap 'k_l2_rerank' 1 k_l2_rerank python tools/l2_time.py
  cap 'k_l2_tc' 1 k_l2_tc python tools/l2_time.py
}
python tools/ncu_traffic.py gpurun_out/prof_k_knn2_tc4_$tag.ncu-rep 1024 5000 1024 > gpurun_out/ncu_traffic_$tag.log 2>&1
# stall samples by source line of the matcher (the drain loop): top lines
ncu -i gpurun_out/prof_k_knn2_tc4_$tag.ncu-rep --page source --csv > gpurun_out/${tag}_tc4_source.csv 2>/dev/null
ls -la gpurun_out | grep $tag | head -40
