tag=r2al
B="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_${tag}_short.log 2>&1 || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
cap() {
  local k=$1 s=$2 o=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s $s -c 1 -o gpurun_out/prof_${o}_$tag -f "$@" > gpurun_out/ncu_${o}_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${o}_$tag.ncu-rep > gpurun_out/${tag}_ncu_$o.txt 2>&1
}
cap 'k_solve8$' 4 k_solve8 $B
grep -E "time_duration|registers_per_thread |fp64|issue_active|warps_active" gpurun_out/${tag}_ncu_k_solve8.txt
