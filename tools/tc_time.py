"""Time the Hamming stage of vb_pairs_run_d on a short synthetic sequence (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
from vslam_b200.lib import Context
ctx = Context(0)
nf, k = 257, 5000
rng = np.random.default_rng(1)
desc = torch.from_numpy(rng.integers(0, 256, (nf, k, 32), dtype=np.uint8)).cuda()
pts = torch.from_numpy((rng.random((nf, k, 2)) * 700).astype(np.float32)).cuda()
import ctypes as C
from vslam_b200 import lib as vl
L = vl.load_library()
prm = ctx.params(0.7, 8, 64, 10.0, 1)
res = torch.zeros((nf - 1) * 64, dtype=torch.uint8, device="cuda")
ctx.profile(True)
for it in range(3):
    rc = L.vb_pairs_run_d(ctx.h, C.c_void_p(pts.data_ptr()), C.c_void_p(desc.data_ptr()), nf, k, 32, C.byref(prm),
                          C.c_void_p(res.data_ptr()), None)
    ctx.synchronize()
    print("rc", rc, "hamming ms", ctx.profile_ms("hamming"), "expand", ctx.profile_ms("expand"), "per pair us",
          1e3 * ctx.profile_ms("hamming") / (nf - 1), flush=True)
