"""Bounded counting in the pair pipeline (k_bq_init / k_count_queue): hypotheses that provably cannot reach the largest
inlier count are abandoned early. The result of find_fundamental (src/RansacFilter.cpp:36-67: winner, inlier count, score, F, mask
and the compacted matches) must not change by one bit — against the same call with the bound switched off, and against the
CPU oracle. Option ransac_prune = 2 forces the bounded path for every batch size (by default it runs for batches that fill the
machine), prune_rounds / prune_growth16 move the checkpoints, prune_item_chunks sizes the queue's work items (vb_set_option)."""
import numpy as np
import pytest

from vslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from vslam_b200.lib import Context
    c = Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same_results(a, ma, b, mb):
    for k in ("status", "n_tentative", "n_matches", "best_hyp", "n_inliers"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(bits(a["score"]), bits(b["score"]))
    assert np.array_equal(bits(a["F"]), bits(b["F"]))
    for i, n in enumerate(a["n_matches"]):
        assert np.array_equal(ma[i, :n], mb[i, :n]), i


CASES = [
    # nframes, k, seed, noise_px, outlier_frac, iters
    (7, 1500, 3, 0.0, 0.2, 256),     # noise-free: many hypotheses tie at the largest count
    (7, 2500, 4, 0.5, 0.3, 1024),    # the benchmark's regime
    (6, 2500, 5, 1.5, 0.7, 512),     # few inliers: the bound is weak, checkpoints late
    (9, 300, 6, 0.5, 0.3, 256),      # fewer matches than the first round covers
    (9, 420, 7, 0.5, 0.1, 300),      # a little more than the first round; hypothesis count not a multiple of 256
    (5, 40, 8, 0.5, 0.3, 64),        # a handful of matches, some pairs below min_items
    (4, 5000, 9, 0.3, 0.05, 1024),   # almost everything an inlier: tiny deficit, many checkpoints
]


@pytest.mark.parametrize("nframes,k,seed,noise,outl,iters", CASES)
def test_bounded_counting_equals_full_counting(ctx, nframes, k, seed, noise, outl, iters):
    pts, desc = synth.sequence(nframes, k, seed, noise_px=noise, outlier_frac=outl)
    prm = ctx.params(0.7, 8, iters, 10.0, 11 + seed)
    ctx.set_option("ransac_prune", 0)
    ref, mref = ctx.pairs_run(pts, desc, prm)
    ctx.set_option("ransac_prune", 2)
    ctx.ransac_prune_stats(reset=True)
    got, mgot = ctx.pairs_run(pts, desc, prm)
    ev, tot = ctx.ransac_prune_stats()
    assert_same_results(got, mgot, ref, mref)
    assert 0 < ev <= tot or tot == 0


@pytest.mark.parametrize("rounds,growth16,item_chunks", [(2, 8, 1), (3, 1, 2), (5, 64, 1), (16, 2, 3), (12, 8, 64)])
def test_bounded_counting_checkpoint_schedules(ctx, rounds, growth16, item_chunks):
    pts, desc = synth.sequence(5, 3000, 21, noise_px=0.5, outlier_frac=0.3)
    prm = ctx.params(0.7, 8, 512, 10.0, 3)
    ctx.set_option("ransac_prune", 0)
    ref, mref = ctx.pairs_run(pts, desc, prm)
    ctx.set_option("ransac_prune", 2)
    ctx.set_option("prune_rounds", rounds)
    ctx.set_option("prune_growth16", growth16)
    ctx.set_option("prune_item_chunks", item_chunks)
    got, mgot = ctx.pairs_run(pts, desc, prm)
    assert_same_results(got, mgot, ref, mref)


def test_bounded_counting_saves_work_and_matches_oracle(ctx, oracle):
    pts, desc = synth.sequence(4, 2500, 31, noise_px=0.5, outlier_frac=0.3)
    prm = ctx.params(0.7, 8, 1024, 10.0, 500)
    ctx.set_option("ransac_prune", 2)
    ctx.ransac_prune_stats(reset=True)
    res, mm = ctx.pairs_run(pts, desc, prm)
    ev, tot = ctx.ransac_prune_stats()
    assert tot == int(sum(1024 * int(t) for t in res["n_tentative"]))
    assert ev < 0.8 * tot, (ev, tot)   # an inlier-rich sequence: most hypotheses come from contaminated samples and die early
    for i in range(3):
        o = oracle.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, 1024, 10.0, 500 + i)
        assert res["n_matches"][i] == o["n"] and res["best_hyp"][i] == o["best"] and res["n_tentative"][i] == o["n_tentative"]
        assert np.array_equal(mm[i, :o["n"]], o["matches"])
        assert np.array_equal(bits(res["F"][i]).reshape(-1), bits(o["F"]).reshape(-1))


def test_bounded_counting_single_problem_entry_point(ctx, oracle):
    """vb_ransac_fundamental (one problem) through the bounded path, including a threshold of 0 (no inliers anywhere: nothing
    can be abandoned) and random correspondences (no model stands out)."""
    fp = synth.frame_pair(1200, 41, outlier_frac=0.4)
    tent = oracle.match_hamming(fp["d1"], fp["d2"], 0.7)
    rng = np.random.default_rng(2)
    junk1 = (rng.random((900, 2)) * 600).astype(np.float32)
    junk2 = (rng.random((900, 2)) * 600).astype(np.float32)
    jm = np.stack([np.arange(900), rng.permutation(900)], 1).astype(np.int32)
    for p1, p2, m, thr in ((fp["p1"], fp["p2"], tent, 10.0), (fp["p1"], fp["p2"], tent, 0.0), (junk1, junk2, jm, 10.0),
                           (fp["p1"], fp["p2"], tent, 1e30)):
        ctx.set_option("ransac_prune", 2)
        g = ctx.ransac_fundamental(p1, p2, m, 8, 600, thr, 17)
        o = oracle.find_fundamental(p1, p2, m, 8, 600, thr, 17)
        assert (g["rc"] == 5) == (o["best"] < 0)
        assert g["best"] == o["best"] and g["n_inliers"] == o["n_inliers"]
        if o["best"] < 0:
            continue
        assert np.array_equal(bits(g["score"]), bits(o["score"]))
        assert np.array_equal(bits(g["F"]), bits(o["F"]))
        assert np.array_equal(g["mask"], o["mask"])


def test_bounded_counting_default_rule_device_batch(ctx):
    """vb_pairs_run_d on a device-resident batch large enough for the default rule (option ransac_prune unset) and for the two
    batch halves on two streams: same results as the full count; and duplicated / degenerate correspondences in the batch
    (every hypothesis ties, or no hypothesis has an inlier) do not disturb the queue."""
    import ctypes as C
    import torch
    from vslam_b200.lib import PAIR_RESULT_DTYPE
    nframes, k = 701, 400
    pts, desc = synth.sequence(nframes, k, 77, noise_px=0.5, outlier_frac=0.3)
    pts[100:103] = pts[100]                     # identical frames: noise-free duplicates, every hypothesis ties
    desc[100:103] = desc[100]
    pts[300] = 5.0                              # all keypoints of a frame in one place: degenerate models
    prm = ctx.params(0.7, 8, 512, 10.0, 5)
    P = nframes - 1
    pts_d, desc_d = torch.from_numpy(pts).cuda(), torch.from_numpy(desc).cuda()

    def run():
        res = torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        out = torch.zeros((P, k, 2), dtype=torch.int32, device="cuda")
        ctx._chk(ctx.L.vb_pairs_run_d(ctx.h, C.c_void_p(pts_d.data_ptr()), C.c_void_p(desc_d.data_ptr()), nframes, k, 32,
                                      C.byref(prm), C.c_void_p(res.data_ptr()), C.c_void_p(out.data_ptr())))
        ctx.synchronize()
        return res.cpu().numpy().view(PAIR_RESULT_DTYPE), out.cpu().numpy()

    ctx.set_option("ransac_prune", 0)
    ref, mref = run()
    ctx.reset_options()
    ctx.ransac_prune_stats(reset=True)
    got, mgot = run()
    ev, tot = ctx.ransac_prune_stats()
    assert 0 < ev < tot            # the default rule took the bounded path and abandoned something
    assert_same_results(got, mgot, ref, mref)
    assert (got["status"] == 0).sum() > 600
