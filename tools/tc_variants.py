"""Correctness + time of the tcgen05 Hamming matcher's drain variants (option tc_drain) on the GPU box (development aid).
Prints one line per variant: ok flag, ms per 256-pair launch, us per 5k x 5k pair."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
from oracle_lib import Oracle
from vslam_b200.lib import Context
ctx, orc = Context(0), Oracle()
rng = np.random.default_rng(5)
nf, k = 257, 5000
desc = torch.from_numpy(rng.integers(0, 256, (nf, k, 32), dtype=np.uint8)).cuda()
pts = torch.from_numpy((rng.random((nf, k, 2)) * 700).astype(np.float32)).cuda()
prm = ctx.params(0.7, 8, 64, 10.0, 1)
res = torch.zeros((nf - 1) * 64, dtype=torch.uint8, device="cuda")
shapes = [(5000, 5000), (300, 700), (1000, 513), (255, 241), (2500, 16384)]
cases = []
for n1, n2 in shapes:
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    d2[n2 // 2] = d2[3]; d1[5 % n1] = d2[3]
    cases.append((d1, d2) + orc.knn2_hamming(d1, d2) + (orc.match_hamming(d1, d2, 0.9),))
extra = [kv for kv in os.environ.get("TC_EXTRA", "").split(",") if kv]
for drain in [int(x) for x in os.environ.get("TC_DRAINS", "0,1,2").split(",")]:
    ctx.reset_options()
    ctx.set_option("hamming_tc", 1)
    ctx.set_option("tc_drain", drain)
    for kv in extra:
        n_, v_ = kv.split("="); ctx.set_option(n_, int(v_))
    ok = True
    for d1, d2, oi, od, om in cases:
        idx, dist = ctx.knn2_hamming(d1, d2)   # (variant 6 hands callers that want the second index to variant 1)
        ok &= bool(np.array_equal(idx, oi) and np.array_equal(dist, od))
        ok &= bool(np.array_equal(ctx.match_hamming(d1, d2, 0.9), om))
    ctx.profile(True)
    ms = []
    for it in range(4):
        rc = ctx.L.vb_pairs_run_d(ctx.h, C.c_void_p(pts.data_ptr()), C.c_void_p(desc.data_ptr()), nf, k, 32, C.byref(prm),
                                  C.c_void_p(res.data_ptr()), None)
        ctx.synchronize()
        ms.append(ctx.profile_ms("hamming"))
    ctx.profile(False)
    print(f"tc_drain={drain} {' '.join(extra)} ok={ok} hamming_ms={min(ms[1:]):.4f} us_per_pair={1e3 * min(ms[1:]) / (nf - 1):.3f}", flush=True)
