"""Time the float-descriptor kNN-2 at BASELINE config-3 size (20000 x 20000 x 128) on the GPU (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from vslam_b200.lib import Context
ctx = Context(0)
rng = np.random.default_rng(5)
n, dim = int(os.environ.get("N", 20000)), 128
d2 = rng.standard_normal((n, dim)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
d1 = (d2[rng.permutation(n)] + 0.05 * rng.standard_normal((n, dim))).astype(np.float32)
ctx.profile(True)
for it in range(3):
    idx, dist = ctx.knn2_l2f(d1, d2)
    print({k: round(ctx.profile_ms(k), 4) for k in ("l2f", "l2f_gemm", "l2f_rerank")}, flush=True)
