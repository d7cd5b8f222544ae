"""On-GPU check of the tensor-core float-descriptor path against the CPU oracle + timing (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from oracle_lib import Oracle
from vslam_b200.lib import Context

ctx, orc = Context(0), Oracle()
ctx.set_option("l2_tc", 1)
rng = np.random.default_rng(5)
ok = True
for n1, n2, dim, kind in [(256, 256, 128, "unit"), (300, 700, 64, "unit"), (1000, 513, 128, "gauss"), (2000, 3000, 128, "sift"),
                          (37, 2, 128, "unit"), (500, 40, 64, "dup"), (4000, 4000, 128, "unit")]:
    d2 = rng.standard_normal((n2, dim)).astype(np.float32)
    d1 = rng.standard_normal((n1, dim)).astype(np.float32)
    if kind == "unit":
        d1 /= np.linalg.norm(d1, axis=1, keepdims=True); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        m = min(n1, n2) // 2
        d1[:m] = d2[:m] + 0.05 * rng.standard_normal((m, dim)).astype(np.float32)
    elif kind == "sift":
        d1 = np.abs(d1 * 40).round().astype(np.float32); d2 = np.abs(d2 * 40).round().astype(np.float32)   # integer-valued: many exact ties
    elif kind == "dup":
        d2[:] = d2[0]; d2[7] += 1e-3                                   # every chunk minimum equal: exact fallback scan
    if n2 > 40:
        d2[n2 // 2] = d2[3]; d1[5] = d2[3]                             # duplicate rows: tie order, zero distance
    t = time.time(); idx, dist = ctx.knn2_l2f(d1, d2); dt = time.time() - t
    oi, od = orc.knn2_l2f(d1, d2)
    good = np.array_equal(idx, oi) and np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    ok &= good
    print(n1, n2, dim, kind, "OK" if good else "MISMATCH", f"{dt*1e3:.2f} ms", flush=True)
    if not good:
        bad = np.nonzero((idx != oi).any(1) | (dist != od).any(1))[0]
        print(" bad rows", bad[:10], "of", len(bad))
        for b in bad[:5]:
            print("  ", b, idx[b], dist[b], oi[b], od[b])
# timing at config-3 size
n, dim = 20000, 128
d2 = rng.standard_normal((n, dim)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
d1 = (d2[rng.permutation(n)] + 0.05 * rng.standard_normal((n, dim))).astype(np.float32)
for tc in ("1", "0"):
    ctx.set_option("l2_tc", int(tc))
    ctx.profile(True)
    for it in range(3):
        idx, dist = ctx.knn2_l2f(d1, d2)
    print("TC" if tc == "1" else "SIMT", {k: round(ctx.profile_ms(k), 4) for k in ("l2f", "l2f_gemm", "l2f_rerank")}, flush=True)
    ctx.profile(False)
    if tc == "1": ref = (idx.copy(), dist.copy())
print("20000 TC == SIMT:", np.array_equal(ref[0], idx) and np.array_equal(ref[1], dist))
sys.exit(0 if ok else 1)
