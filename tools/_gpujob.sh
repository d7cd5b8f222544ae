tag=r2ab
B="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_${tag}_short.log 2>&1 || { echo "bench failed"; tail -5 gpurun_out/bench_${tag}_short.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
cap() {
  local k=$1 s=$2 o=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s $s -c 1 -o gpurun_out/prof_${o}_$tag -f "$@" > gpurun_out/ncu_${o}_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${o}_$tag.ncu-rep > gpurun_out/${tag}_ncu_$o.txt 2>&1
}
cap k_knn2_tc4 4 k_knn2_tc4 $B
cap k_count_queue 4 k_count_queue $B
cap k_knn2_tc_fix 4 k_knn2_tc_fix $B
python tools/ncu_traffic.py gpurun_out/prof_k_knn2_tc4_$tag.ncu-rep 1024 5000 1024 > gpurun_out/ncu_traffic_$tag.log 2>&1
python tools/ncu_traffic.py gpurun_out/prof_k_count_queue_$tag.ncu-rep 1024 5000 1024 >> gpurun_out/ncu_traffic_$tag.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_$tag.json
head -12 gpurun_out/${tag}_ncu_k_knn2_tc4.txt; wc -l gpurun_out/launches_$tag.csv
