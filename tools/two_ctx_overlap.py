"""Experiment: consecutive 1 024-pair steps alternating between two contexts (two streams), so that the small kernels and the
counting of step n may run beside the matcher of step n+1. Prints ms per step for one context and for two."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
from vslam_b200 import synth
from vslam_b200.lib import PAIR_RESULT_DTYPE, Context
dev = torch.device("cuda", 0)
nframes, k = 1025, 5000
P = nframes - 1
pts, desc = synth.sequence(nframes, k, 1000)
pts_d, desc_d = torch.from_numpy(pts).to(dev), torch.from_numpy(desc).to(dev)
ctxs = [Context(0), Context(0)]
for c in ctxs:
    for kv in filter(None, os.environ.get("EXP_OPTS", "").split(",")):
        n_, v_ = kv.split("="); c.set_option(n_, int(v_))
prm = ctxs[0].params(0.7, 8, 1024, 10.0, 1)
res = [torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in range(2)]
out = [torch.zeros((P, k, 2), dtype=torch.int32, device=dev) for _ in range(2)]
def step(i, two):
    j = (i & 1) if two else 0
    c = ctxs[j]
    c._chk(c.L.vb_pairs_run_d(c.h, pts_d.data_ptr(), desc_d.data_ptr(), nframes, k, 32, C.byref(prm), res[j].data_ptr(), out[j].data_ptr()))
for two in (False, True, False, True):
    for i in range(4): step(i, two)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for i in range(n): step(i, two)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    r0, r1 = res[0].cpu().numpy().view(PAIR_RESULT_DTYPE), res[1].cpu().numpy().view(PAIR_RESULT_DTYPE)
    print(f"{'two contexts' if two else 'one context '}: {dt:.3f} ms per step, {P / dt:.1f} k pairs/s, results equal: {bool(np.array_equal(r0['n_matches'], r1['n_matches']))}", flush=True)
