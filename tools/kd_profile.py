"""KD-tree throughput batches on their own (the `kdtree` object of bench.py) — the command the ncu captures of k_kd_build /
k_kd_nearest / k_kd_radius profile:  ncu --set full -k regex:k_kd_ ... python tools/kd_profile.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from vslam_b200 import synth
from vslam_b200.lib import Context
ctx = Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
pts, _ = synth.sequence(8, 5000, 1000)
print(json.dumps(bench.kdtree_stage(ctx, torch, dev, pts, 5000, nq_big=int(os.environ.get("NQ", 1 << 20)))))
