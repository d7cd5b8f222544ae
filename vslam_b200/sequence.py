"""Sharding a frame sequence across GPUs.

Pair i = (frame i, frame i+1) depends only on those two frames, so the npairs = nframes-1 pairs of a
sequence are split into contiguous ranges, one per rank; each rank needs its frames plus a ONE-FRAME
HALO (the first frame of the next rank's range). There is no data-path collective: every rank runs
the batched pair pipeline on its own GPU and the per-pair results are gathered on the host (rank 0).
NCCL is deliberately not used.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def shard_pairs(npairs: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [begin, end) pair ranges, sizes differing by at most one, earlier ranks larger."""
    base, extra = divmod(max(npairs, 0), world)
    out, b = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((b, b + n))
        b += n
    return out


def frame_range(pair_range: tuple[int, int]) -> tuple[int, int]:
    """Frames a rank must hold for its pair range: [begin, end + 1) — the +1 is the halo frame."""
    b, e = pair_range
    return (b, e + 1) if e > b else (b, b)


def run_sharded(pts: np.ndarray, desc: np.ndarray, rank: int, world: int, seed0: int,
                run_pairs: Callable[[np.ndarray, np.ndarray, int], np.ndarray]) -> tuple[tuple[int, int], np.ndarray]:
    """Run this rank's share. run_pairs(pts_shard, desc_shard, seed0_shard) -> structured result array
    of len(shard frames) - 1. Pair i keeps seed seed0 + i regardless of the world size, so the sharded
    result equals the single-GPU result pair for pair."""
    pr = shard_pairs(len(pts) - 1, world)[rank]
    fb, fe = frame_range(pr)
    if fe - fb < 2:
        return pr, None
    return pr, run_pairs(pts[fb:fe], desc[fb:fe], seed0 + pr[0])


def gather_results(local: np.ndarray | None, pair_range: tuple[int, int], npairs: int, world: int,
                   gather: Callable[[object], Sequence[object]] | None = None):
    """Host gather onto every caller that supplies `gather` (torch.distributed.all_gather_object-like);
    with world == 1 it is the identity."""
    items = [(pair_range, local)] if gather is None else list(gather((pair_range, local)))
    out = None
    for (b, e), arr in items:
        if arr is None or e <= b:
            continue
        if out is None:
            out = np.zeros(npairs, arr.dtype)
        out[b:e] = arr
    return out
