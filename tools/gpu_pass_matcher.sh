#!/bin/bash
# GPU pass after a change to the matcher: -m gpu tests, the default bench line, the launch list and one ncu capture of
# k_knn2_tc4 and of the float path's GEMM (each ncu command only after the same command ran clean without it). tag = $1.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { echo "bench failed"; tail -5 gpurun_out/${tag}_bench.err; }
B="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_${tag}_short.log 2>&1 || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench_quick_steps2.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
cap() {  # cap <kernel regex> <skip> <out name> <command...>
  local k=$1 s=$2 o=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s $s -c 1 -o gpurun_out/prof_${o}_$tag -f "$@" > gpurun_out/ncu_${o}_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${o}_$tag.ncu-rep > gpurun_out/${tag}_ncu_$o.txt 2>&1
}
cap k_knn2_tc4 4 k_knn2_tc4 $B
python tools/l2_time.py > /dev/null 2>&1 && cap 'k_l2_tc' 1 k_l2_tc python tools/l2_time.py
python tools/ncu_traffic.py gpurun_out/prof_k_knn2_tc4_$tag.ncu-rep 1024 5000 1024 > gpurun_out/ncu_traffic_$tag.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/${tag}_ncu_traffic.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'kernel_ms', d.get('kernel_ms'))"
