"""Host-side model of the bounded counting (vslam_b200/csrc/ransac.cu: k_bq_init / bq_prune / k_count_queue), run against the
CPU oracle's per-hypothesis inlier masks (src/RansacFilter.cpp:105-140 through oracle/vb_oracle.c). It restates the rule —
rounds over chunks of 128 matches, a hypothesis dropped when `count so far + matches not seen < L`, L raised by partial counts
and by the leader's complete count, the first leader's outliers visited first — in numpy and checks the three properties the
GPU path relies on:

  * no hypothesis that reaches the largest inlier count is ever dropped (so find_fundamental's selection, :59, is unchanged);
  * a dropped hypothesis keeps a partial count strictly below the maximum (it can never look tied);
  * the work actually shrinks on an inlier-rich problem, and degenerate problems (no inliers at all / everything tied) keep
    every hypothesis to the end.

The GPU implementation itself is compared bit for bit with the full count in tests/test_gpu_bounded_count.py."""
import numpy as np
import pytest

from vslam_b200 import synth

CHUNK = 128


def inlier_matrix(oracle, fp, tent, iters, thr, seed):
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, iters, thr, seed, want_all=True)
    masks = np.stack([oracle.residual(fp["p1"], fp["p2"], tent, o["F_all"][h].reshape(3, 3), thr)[0] for h in range(iters)])
    assert np.array_equal(masks.sum(1), o["cnt_all"])
    return masks.astype(np.int64), o


def bounded_count(masks, first_chunks=2, first16=20, growth16=6, max_rounds=8):
    """Returns (counts as the device would leave them, alive mask at the end, evaluations done)."""
    H, m = masks.shape
    nchunks = -(-m // CHUNK)
    order = np.arange(m)
    alive = np.arange(H)
    cnt = np.zeros(H, np.int64)
    L, lo, hi, rnd, boosted, evals = 0, 0, min(first_chunks, nchunks), 0, -1, 0
    while True:
        seg = order[lo * CHUNK:min(hi * CHUNK, m)]
        cnt[alive] += masks[np.ix_(alive, seg)].sum(1)
        evals += len(alive) * len(seg)
        if hi >= nchunks:
            return cnt, alive, evals
        remaining = m - hi * CHUNK
        leader = alive[np.lexsort((alive, -cnt[alive]))[0]]          # largest count so far, lowest index
        L = max(L, int(cnt[leader]))
        if rnd == 0:
            seen = hi * CHUNK
            rest = np.arange(seen, m)
            out_first = np.concatenate([rest[masks[leader, rest] == 0], rest[masks[leader, rest] == 1][::-1]])
            order = np.concatenate([np.arange(seen), out_first])     # leader's outliers first (inliers' order is irrelevant)
            L = max(L, int(masks[leader].sum()))
        elif leader != boosted:
            L = max(L, int(masks[leader].sum()))
        boosted = leader
        alive = alive[cnt[alive] + remaining >= L]
        assert len(alive) >= 1
        done = hi
        if rnd + 2 >= max_rounds:
            nxt = nchunks
        elif rnd == 0:
            nxt = -(-(first16 * (m - L) // 16) // CHUNK)
        else:
            nxt = done + max(1, done * growth16 // 16)
        lo, hi, rnd = done, min(max(nxt, done + 1), nchunks), rnd + 1


@pytest.mark.parametrize("k,outl,iters,thr,seed", [(1500, 0.3, 200, 10.0, 1), (1200, 0.7, 150, 10.0, 2), (900, 0.1, 120, 10.0, 3),
                                                   (700, 0.3, 100, 0.0, 4), (300, 0.3, 64, 10.0, 5)])
def test_bounded_counting_never_drops_a_maximal_hypothesis(oracle, k, outl, iters, thr, seed):
    fp = synth.frame_pair(k, seed, outlier_frac=outl)
    tent = oracle.match_hamming(fp["d1"], fp["d2"], 0.7)
    masks, o = inlier_matrix(oracle, fp, tent, iters, thr, 50 + seed)
    full = masks.sum(1)
    for sched in ((2, 20, 6, 8), (1, 16, 2, 16), (3, 40, 16, 3), (2, 20, 6, 2)):
        cnt, alive, evals = bounded_count(masks, *sched)
        top = np.nonzero(full == full.max())[0]
        assert set(top) <= set(alive), sched                                     # every maximal hypothesis survives
        assert np.array_equal(cnt[alive], full[alive]), sched                    # survivors carry complete counts
        dropped = np.setdiff1d(np.arange(len(full)), alive)
        assert (cnt[dropped] < full.max()).all() and (cnt[dropped] <= full[dropped]).all(), sched
        assert evals <= masks.size
        if thr == 0.0:
            assert len(alive) == len(full)                                       # nothing has an inlier: nothing can be dropped
    if outl <= 0.3 and thr > 0 and len(tent) > 4 * CHUNK:
        _, _, evals = bounded_count(masks)
        assert evals < 0.7 * masks.size                                          # the rule does save work where it should


def test_bounded_counting_all_tied(oracle):
    fp = synth.frame_pair(400, 9, noise_px=0.0, outlier_frac=0.0)
    tent = np.stack([np.arange(400), fp["gt"]], 1).astype(np.int32)
    masks, o = inlier_matrix(oracle, fp, tent, 64, 10.0, 77)
    full = masks.sum(1)
    cnt, alive, _ = bounded_count(masks)
    assert set(np.nonzero(full == full.max())[0]) <= set(alive)
    assert np.array_equal(cnt[alive], full[alive])
