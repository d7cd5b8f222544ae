python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2n_gpu_tests.log; cat gpurun_out/r2n_gpu_tests.log
cap() { local k=$1 s=$2 o=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s $s -c 1 -o gpurun_out/prof_${o}_r2n -f "$@" > gpurun_out/ncu_${o}_r2n.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${o}_r2n.ncu-rep > gpurun_out/r2n_ncu_$o.txt 2>&1; grep -E "gpu__time_duration|grid_size" gpurun_out/r2n_ncu_$o.txt | head -3; }
# the 2^20-query launches: skip the latency-sized launches of part (1) of kdtree_stage (8 of each kernel) and the warm-ups
cap 'k_kd_nearest<\(int\)1>' 10 k_kd_nearest_1M python tools/kd_profile.py
cap 'k_kd_radius<\(bool\)0>' 10 k_kd_radius_1M python tools/kd_profile.py
cap 'k_kd_knn' 0 k_kd_knn python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np
from vslam_b200.lib import Context
ctx=Context(0); rng=np.random.default_rng(0)
pts=np.stack([rng.uniform(0,1280,5000),rng.uniform(0,720,5000)],1).astype(np.float32)
t=ctx.kdtree_build(pts); q=(pts[rng.integers(0,5000,1<<20)]+rng.uniform(-3,3,(1<<20,2))).astype(np.float32)
t.knn(q,8)"
(time python bench.py) > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -4 gpurun_out/r2n_bench.err
