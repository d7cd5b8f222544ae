"""Timeline of one CTA of the tcgen05 Hamming matcher (development aid, TUNING build: VB_LIB_PATH=vslam_b200/lib_tuning/...).
For each drain variant in TC_DRAINS: runs a 256-pair launch with tc_dbg = 32 and prints, per step (= 2 * tile + accumulator)
of CTA 0, the SM-clock offsets of the hand-shake points of the UMMA thread, two draining warps and a re-arming warp, followed
by the averages over the steady-state steps."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from vslam_b200.lib import Context
ctx = Context(0)
rng = np.random.default_rng(5)
nf, k = 257, 5000
desc = torch.from_numpy(rng.integers(0, 256, (nf, k, 32), dtype=np.uint8)).cuda()
pts = torch.from_numpy((rng.random((nf, k, 2)) * 700).astype(np.float32)).cuda()
prm = ctx.params(0.7, 8, 64, 10.0, 1)
res = torch.zeros((nf - 1) * 64, dtype=torch.uint8, device="cuda")
S, E, R = 128, 6, 18   # steps, events, roles (0 = UMMA thread, 1 = re-arming warp 20, 2 + ew = draining warp ew)
ctx.L.vb_debug_tc_trace.argtypes = [C.c_void_p, C.c_int]
for drain in [int(x) for x in os.environ.get("TC_DRAINS", "6,8,9").split(",")]:
    ctx.reset_options()
    ctx.set_option("hamming_tc", 1)
    ctx.set_option("tc_drain", drain)
    ctx.set_option("tc_dbg", 32)
    for kv in filter(None, os.environ.get("TC_EXTRA", "").split(",")):
        n_, v_ = kv.split("="); ctx.set_option(n_, int(v_))
    for it in range(2):
        ctx.L.vb_pairs_run_d(ctx.h, C.c_void_p(pts.data_ptr()), C.c_void_p(desc.data_ptr()), nf, k, 32, C.byref(prm),
                             C.c_void_p(res.data_ptr()), None)
        ctx.synchronize()
    buf = np.zeros(R * S * E, dtype=np.int64)
    assert ctx.L.vb_debug_tc_trace(buf.ctypes.data, buf.size) == 0
    raw = buf.reshape(R, S, E)
    polls = raw[0, :, 5].copy()
    t = raw.astype(np.float64)
    t0 = t[0, 0, 0]
    t = np.where(t > 0, t - t0, np.nan)
    rearm = not np.isnan(t[1, 8, 1])
    dr = t[2:18]                                     # [16 warps][steps][events]
    full_first, full_last = np.nanmin(dr[:, :, 1], 0), np.nanmax(dr[:, :, 1], 0)
    ld_last = np.nanmax(dr[:, :, 2], 0)
    hb_last = t[1, :, 2] if rearm else np.nanmax(dr[:, :, 3], 0)     # accumulator re-armed (all warps / the traced re-arming warp)
    alu_last = np.nanmax(dr[:, :, 4], 0)
    print(f"== tc_drain={drain}: step | UMMA thread: wait_tempty(polls) passed commit_issued | tfull seen first..last | loads done (last warp) | handed back (last) | max trees done (last)")
    for s in range(8, 32):
        m = t[0, s]
        print(f"{s:3d} | {m[0]:7.0f} ({polls[s]:2d}) {m[1]:7.0f} {m[2]:7.0f} | {full_first[s]:7.0f} {full_last[s]:7.0f} | {ld_last[s]:7.0f} | {hb_last[s]:7.0f} | {alu_last[s]:7.0f}")
    lo, hi = 8, 40   # inside the first unit (42 steps), past the ramp
    sl = slice(lo, hi)
    avg = lambda x: float(np.nanmean(x))
    print(f"   clk per step {(t[0, hi, 1] - t[0, lo, 1]) / (hi - lo):.0f}")
    print(f"   UMMA thread: in the tempty wait {avg(t[0, sl, 1] - t[0, sl, 0]):.0f} (failed polls per step {polls[sl].mean():.2f}), issue of 4 UMMAs + commit {avg(t[0, sl, 2] - t[0, sl, 1]):.0f}")
    print(f"   tempty passed -> tfull seen by the first warp {avg(full_first[sl] - t[0, sl, 1]):.0f}; commit issued -> tfull seen {avg(full_first[sl] - t[0, sl, 2]):.0f}; first..last warp {avg(full_last[sl] - full_first[sl]):.0f}")
    print(f"   tfull seen (first) -> all loads done {avg(ld_last[sl] - full_first[sl]):.0f} -> handed back {avg(hb_last[sl] - full_first[sl]):.0f} -> max trees done {avg(alu_last[sl] - full_first[sl]):.0f}")
    nxt = t[0, lo + 2:hi + 2, 1]   # the UMMA thread passes the tempty wait of step s + 2 (same accumulator)
    print(f"   handed back (step s) -> UMMA thread passes tempty of step s+2: {avg(nxt - hb_last[sl]):.0f}; it started waiting {avg(t[0, lo + 2:hi + 2, 0] - hb_last[sl]):.0f} after the hand-back (negative = it was waiting)")
    print(f"   per draining warp and step: waiting for tfull {avg(dr[:, sl, 1] - dr[:, sl, 0]):.0f}, loads {avg(dr[:, sl, 2] - dr[:, sl, 1]):.0f}, to hand-back/arrive {avg(dr[:, sl, 3] - dr[:, sl, 2]):.0f}, max trees {avg(dr[:, sl, 4] - dr[:, sl, 3]):.0f}", flush=True)
