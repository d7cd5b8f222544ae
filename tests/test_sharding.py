"""Multi-GPU host logic on CPU: contiguous pair ranges with a one-frame halo, per-pair seeds
independent of the world size, host gather. Run with world_size 2 over gloo; the per-shard compute is
injected (here: the oracle), the product's GPU path plugs in through the same callable."""
import os
import subprocess
import sys

import numpy as np
import pytest

from vslam_b200 import sequence as seq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_pairs_partition():
    for npairs in (0, 1, 2, 7, 8, 9999):
        for world in (1, 2, 3, 4, 8):
            sh = seq.shard_pairs(npairs, world)
            assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == npairs
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [e - b for b, e in sh]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            for pr in sh:
                fb, fe = seq.frame_range(pr)
                assert fe - fb == (pr[1] - pr[0] + 1 if pr[1] > pr[0] else 0)      # +1 = halo frame


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from oracle_lib import Oracle
from vslam_b200 import sequence as seq, synth
from vslam_b200.lib import PAIR_RESULT_DTYPE
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
orc = Oracle()
pts, desc = synth.sequence(8, 300, 5)          # identical on every rank (same seed)

def run_pairs(p, d, seed0):
    out = np.zeros(len(p) - 1, PAIR_RESULT_DTYPE)
    for i in range(len(p) - 1):
        o = orc.match_features(p[i], d[i], p[i + 1], d[i + 1], 0.7, 8, 32, 10.0, seed0 + i)
        out["status"][i] = 0 if o["n"] >= 0 else 3
        out["n_matches"][i] = max(o["n"], 0); out["n_tentative"][i] = o["n_tentative"]; out["best_hyp"][i] = o["best"]
        out["F"][i] = o["F"].reshape(-1)
    return out

pr, local = seq.run_sharded(pts, desc, rank, world, 77, run_pairs)
def gather(item):
    box = [None] * world
    dist.all_gather_object(box, item)
    return box
full = seq.gather_results(local, pr, len(pts) - 1, world, gather)
if rank == 0:
    np.save({out!r}, full)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_equals_single(tmp_path, oracle):
    from vslam_b200 import synth
    from vslam_b200.lib import PAIR_RESULT_DTYPE
    out = str(tmp_path / "gathered.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = np.load(out)
    pts, desc = synth.sequence(8, 300, 5)
    assert got.dtype == PAIR_RESULT_DTYPE and len(got) == 7
    for i in range(7):      # pair i keeps seed 77 + i whichever rank ran it
        o = oracle.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, 32, 10.0, 77 + i)
        assert got["n_matches"][i] == max(o["n"], 0) and got["best_hyp"][i] == o["best"]
        assert np.array_equal(got["F"][i].view(np.uint32), o["F"].reshape(-1).view(np.uint32))


def test_reflected_sequence_is_made_of_genuine_pairs():
    """bench.py's config-4 sequence (10 001 frames out of a 1 025-frame base walked forwards and backwards): every
    consecutive pair of the long sequence is a consecutive pair of the base, whatever range a rank takes."""
    import bench
    period = 1024
    full = bench.reflected_frames(0, 10001, period)
    assert full.min() == 0 and full.max() == period
    assert (np.abs(np.diff(full)) == 1).all()
    for world in (1, 2, 3, 8):
        from vslam_b200.sequence import shard_pairs
        got = []
        for r in range(world):
            b, e = shard_pairs(10000, world)[r]
            idx = bench.reflected_frames(b, e - b + 1, period)          # the rank's frames incl. the one-frame halo
            assert np.array_equal(idx, full[b:e + 1])
            got.append(e - b)
        assert sum(got) == 10000


def test_multi_partition_matches_header_contract():
    """vb_multi cuts P pairs into ranges [P*w/n, P*(w+1)/n) (multi.cu); ranges tile [0, P) without gaps for any n, and a
    range's packed matches start at first * k (include/vslam_b200.h)."""
    for P in (0, 1, 5, 1024, 10000):
        for n in (1, 2, 3, 8):
            firsts = [P * w // n for w in range(n + 1)]
            assert firsts[0] == 0 and firsts[-1] == P and all(a <= b for a, b in zip(firsts, firsts[1:]))
