"""UMMA-only rate of tcgen05.mma kind::mxf4 (M128, K64) against the tile width N on this GPU (vb_probe_tensor_peak):
what a matcher with narrower train tiles — e.g. three 160-column accumulators — could reach at best."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from vslam_b200.lib import Context
ctx = Context(0)
for n in (256, 240, 208, 192, 176, 160, 144, 128, 112, 96, 80, 64):
    tf = ctx.probe_tensor_peak(0, n)
    clk = 2.0 * 128 * n * 64 / (tf * 1e12 / 148 / 1.965e9)
    print(f"mxf4 M128 N{n:3d} K64: {tf:8.1f} TFLOP/s, {clk:6.1f} cycles per UMMA at 1.965 GHz, "
          f"{(4096 + 32 * n) / clk:5.1f} B of shared memory per cycle", flush=True)
