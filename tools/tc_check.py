"""Quick on-GPU check of the tensor-core Hamming path against the CPU oracle (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from oracle_lib import Oracle
from vslam_b200.lib import Context

ctx, orc = Context(0), Oracle()
ctx.set_option("hamming_tc", 1)
rng = np.random.default_rng(5)
ok = True
for n1, n2 in [(256, 256), (300, 700), (1000, 513), (5000, 5000), (37, 2)]:
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    if n2 > 40:
        d2[n2 // 2] = d2[3]          # duplicate rows: tie order
        d1[5] = d2[3]
    t = time.time()
    idx, dist = ctx.knn2_hamming(d1, d2)
    dt = time.time() - t
    oi, od = orc.knn2_hamming(d1, d2)
    good = np.array_equal(idx, oi) and np.array_equal(dist, od)
    ok &= good
    print(n1, n2, "OK" if good else "MISMATCH", f"{dt*1e3:.2f} ms", flush=True)
    if not good:
        bad = np.nonzero((idx != oi).any(1) | (dist != od).any(1))[0]
        print(" bad rows", bad[:10], "of", len(bad))
        for b in bad[:5]:
            print("  ", b, idx[b], dist[b], oi[b], od[b])
sys.exit(0 if ok else 1)
