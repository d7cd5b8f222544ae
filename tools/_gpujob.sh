export VB_LIB_PATH=$PWD/vslam_b200/lib_tuning/libvslam_b200.so
for o in "tc_drain=1" "tc_drain=1,pairs_split=1" "tc_drain=0,pairs_split=1" "tc_drain=3,pairs_split=1"; do
  VB_OPTIONS=$o python bench.py --quick --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$o', 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done > gpurun_out/r2g_split.log 2>&1
cat gpurun_out/r2g_split.log
