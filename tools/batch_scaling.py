"""Device-resident pairs/s of vb_pairs_run_d as a function of the batch size (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ctypes as C
import numpy as np, torch
from vslam_b200.lib import Context, PAIR_RESULT_DTYPE
from vslam_b200 import synth
ctx = Context(0)
k = 5000
pts, desc = synth.sequence(1025, k, 1000)
pts_d, desc_d = torch.from_numpy(pts).cuda(), torch.from_numpy(desc).cuda()
prm = ctx.params(0.7, 8, 1024, 10.0, 1)
for P in (32, 64, 128, 192, 256, 512, 1024):
    res = torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    out = torch.zeros((P, k, 2), dtype=torch.int32, device="cuda")
    run = lambda: ctx._chk(ctx.L.vb_pairs_run_d(ctx.h, pts_d.data_ptr(), desc_d.data_ptr(), P + 1, k, 32, C.byref(prm),
                                               res.data_ptr(), out.data_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(3, 2048 // P)
    for _ in range(reps): run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    ctx.profile(True); run(); torch.cuda.synchronize()
    km = {n: round(ctx.profile_ms(n) * 1e3 / P, 2) for n in ("expand", "hamming", "knnfix", "finish", "sample", "solve", "score", "select")}
    ctx.profile(False)
    print(P, f"{dt*1e3:.3f} ms  {dt*1e6/P:.2f} us/pair", km, flush=True)
