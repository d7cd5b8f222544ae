python -m pytest tests -x -q -m gpu -k "kdtree or projection or adapter" 2>&1 | tail -15 > gpurun_out/r2i_kd_tests.log; cat gpurun_out/r2i_kd_tests.log
NQ=1048576 python tools/kd_profile.py > gpurun_out/r2i_kd.json 2> gpurun_out/r2i_kd.err; tail -3 gpurun_out/r2i_kd.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_kd.json'))
for r in d['throughput']: print(r['tree_points'], 'build_ms', r['build_ms_single_tree'], 'nearest_ms', r['nearest']['ms'], 'radius_ms', r['radius2']['ms'])"
