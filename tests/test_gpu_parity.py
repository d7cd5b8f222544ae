"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs. Bit-exact for indices, masks, counts, scores and F (the solve is a fully specified op
sequence on both sides); the documented tolerance for F against anything else is 1e-5 relative
Frobenius after sign/norm normalisation (BASELINE.json north_star)."""
import ctypes as C

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


def rel_frob(a, b):
    a, b = a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


# ---------------------------------------------------------------- 8-point solve
def test_solve8_bit_exact(ctx, oracle):
    fp = synth.frame_pair(800, 21)
    rng = np.random.default_rng(0)
    gt = fp["gt"]
    keep = np.nonzero(gt >= 0)[0]
    H = 500
    sel = np.stack([rng.choice(keep, 8, replace=False) for _ in range(H)])
    p1s, p2s = fp["p1"][sel], fp["p2"][gt[sel]]
    # degenerate sets too: repeated point, collinear points, all-zero
    p1s[0] = p1s[0, 0]; p2s[0] = p2s[0, 0]
    p1s[1, :, 1] = 100.0; p2s[1, :, 1] = 100.0
    p1s[2] = 0; p2s[2] = 0
    F = ctx.ransac_solve8(p1s, p2s)
    for h in range(H):
        Fo = oracle.compute_fundamental(p1s[h], p2s[h])
        assert np.array_equal(bits(F[h]), bits(Fo)), h
        if h > 2:
            assert rel_frob(F[h], Fo) <= 1e-5


def test_solve8_golden_inputs(ctx, oracle, golden):
    F = ctx.ransac_solve8(golden["fm_p1"], golden["fm_p2"])
    for h in range(len(F)):
        assert np.array_equal(bits(F[h]), bits(oracle.compute_fundamental(golden["fm_p1"][h], golden["fm_p2"][h])))
        assert rel_frob(F[h], golden["fm_F2"][h]) < 5e-3      # cv2's fp32 SVD is the noisy side


# ---------------------------------------------------------------- scoring kernel
def _oracle_scores(oracle, corr, Fs, thr):
    m = len(corr)
    p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
    matches = np.stack([np.arange(m), np.arange(m)], 1).astype(np.int32)
    cnt, sc = [], []
    for F in Fs:
        _, _, n, s = oracle.residual(p1, p2, matches, F, thr)
        cnt.append(n); sc.append(s)
    return np.array(cnt, np.int32), np.array(sc, np.float32)


def _hyp_bank(oracle, corr, H, seed):
    rng = np.random.default_rng(seed)
    Fs = np.zeros((H, 9), np.float32)
    for h in range(H):
        sel = rng.choice(len(corr), 8, replace=len(corr) < 8)
        Fs[h] = oracle.compute_fundamental(corr[sel, :2], corr[sel, 2:]).reshape(-1)
    return Fs


@pytest.mark.parametrize("m", [1, 7, 127, 128, 129, 1000, 8191, 8192, 8193, 20000])
def test_score_bit_exact_sizes(ctx, oracle, m):
    corr = synth.correspondences(max(m, 8), m)[:m].copy()
    H = 70
    Fs = _hyp_bank(oracle, synth.correspondences(64, 1), H, m)
    cnt, sc = ctx.ransac_score(corr, Fs, 10.0)
    ocnt, osc = _oracle_scores(oracle, corr, Fs, 10.0)
    assert np.array_equal(cnt, ocnt)
    assert np.array_equal(bits(sc), bits(osc))


@pytest.mark.parametrize("H", [1, 127, 129, 257, 600])
def test_score_bit_exact_hypothesis_counts(ctx, oracle, H):
    corr = synth.correspondences(3000, 5)
    Fs = _hyp_bank(oracle, corr, H, H)
    cnt, sc = ctx.ransac_score(corr, Fs, 10.0)
    ocnt, osc = _oracle_scores(oracle, corr, Fs, 10.0)
    assert np.array_equal(cnt, ocnt) and np.array_equal(bits(sc), bits(osc))
    assert H < 100 or cnt.max() > 100     # realistic hypotheses, not all-outlier noise


def test_score_golden_and_special_values(ctx, oracle, golden):
    g = golden
    p1, p2, mm = g["res_p1"], g["res_p2"], g["res_matches"]
    corr = np.ascontiguousarray(np.concatenate([p1[mm[:, 0]], p2[mm[:, 1]]], 1), np.float32)
    cnt, sc = ctx.ransac_score(corr, g["res_F"].reshape(-1, 9), float(g["res_thr"]))
    assert np.array_equal(cnt, g["res_cnt"])                       # cv2-generated counts
    for i in range(len(cnt)):
        _, _, n, s = oracle.residual(p1, p2, mm, g["res_F"][i], float(g["res_thr"]))
        assert cnt[i] == n
        assert bits(sc[i:i + 1])[0] == bits(np.array([s]))[0] or (np.isnan(sc[i]) and np.isnan(s))
    assert np.isnan(sc).any() and np.isinf(sc).any()


def test_score_large_grouped_path(ctx, oracle):
    """m large enough that the kernel folds whole 64-chunk groups inside a CTA."""
    m, H = 300000, 1024
    corr = synth.correspondences(m, 77)
    Fs = _hyp_bank(oracle, corr[:5000], H, 3)
    cnt, sc = ctx.ransac_score(corr, Fs, 10.0)
    pick = np.r_[0:8, H - 8:H, 500:508]
    ocnt, osc = _oracle_scores(oracle, corr, Fs[pick], 10.0)
    assert np.array_equal(cnt[pick], ocnt) and np.array_equal(bits(sc[pick]), bits(osc))


def test_score_linearity_property_full_size(ctx, oracle):
    """Size-independent property at BASELINE size (1M matches): inlier counts are additive over any split
    of the match set, and permuting hypotheses permutes the outputs."""
    m, H = 1000000, 256
    corr = synth.correspondences(m, 5)
    Fs = _hyp_bank(oracle, corr[:4000], H, 9)
    cnt, sc = ctx.ransac_score(corr, Fs, 10.0)
    cut = 333333
    c1, _ = ctx.ransac_score(corr[:cut].copy(), Fs, 10.0)
    c2, _ = ctx.ransac_score(corr[cut:].copy(), Fs, 10.0)
    assert np.array_equal(cnt, c1 + c2)
    perm = np.random.default_rng(0).permutation(H)
    cp, sp = ctx.ransac_score(corr, Fs[perm], 10.0)
    assert np.array_equal(cp, cnt[perm]) and np.array_equal(bits(sp), bits(sc[perm]))
    ocnt, osc = _oracle_scores(oracle, corr, Fs[:3], 10.0)
    assert np.array_equal(cnt[:3], ocnt) and np.array_equal(bits(sc[:3]), bits(osc))


# ---------------------------------------------------------------- sampling + full RANSAC
@pytest.mark.parametrize("n,mi,iters,seed", [(8, 8, 50, 1), (9, 8, 64, 2), (100, 8, 100, 42), (3000, 8, 1024, 7),
                                             (20, 5, 30, 9), (50, 1, 10, 3), (5000, 8, 4096, 11)])
def test_sample_sets_bit_exact(ctx, oracle, n, mi, iters, seed):
    corr = synth.correspondences(n, seed)
    p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
    matches = np.stack([np.arange(n), np.arange(n)], 1).astype(np.int32)
    g = ctx.ransac_hypotheses(p1, p2, matches, mi, iters, 10.0, seed)
    assert np.array_equal(g["sets"], oracle.initialize_sets(n, mi, iters, seed))


def test_sample_sets_with_rejections(ctx, oracle):
    """~1M matches: libstdc++'s Lemire rejection fires about twice per 8192 draws; the kernel must replay it."""
    n, iters, seed = 1000000, 2048, 123
    corr = synth.correspondences(n, 3)
    p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
    matches = np.stack([np.arange(n), np.arange(n)], 1).astype(np.int32)
    g = ctx.ransac_hypotheses(p1, p2, matches, 8, iters, 10.0, seed)
    o = oracle.initialize_sets(n, 8, iters, seed)
    assert np.array_equal(g["sets"], o)

    # the no-rejection replay of the raw stream would have differed (so the branch was exercised)
    class MT(C.Structure):
        _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]
    st = MT()
    oracle.lib.vbo_mt_seed(C.byref(st), seed)
    oracle.lib.vbo_mt_next.restype = C.c_uint32
    rej = 0
    for i in range(iters * 8):
        rng_ = n - (i % 8)
        low = (oracle.lib.vbo_mt_next(C.byref(st)) * rng_) & 0xFFFFFFFF
        rej += low < rng_ and low < ((1 << 32) - rng_) % rng_
    assert rej > 0


@pytest.mark.parametrize("k,iters,thr,seed", [(400, 100, 10.0, 5), (2000, 100, 10.0, 0), (2000, 1024, 10.0, 1),
                                              (300, 64, 0.2, 3), (5000, 1024, 10.0, 2)])
def test_find_fundamental_bit_exact(ctx, oracle, k, iters, thr, seed):
    fp = synth.frame_pair(k, seed)
    tent = oracle.match_hamming(fp["d1"], fp["d2"])
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, iters, thr, seed + 100, want_all=True)
    g = ctx.ransac_fundamental(fp["p1"], fp["p2"], tent, 8, iters, thr, seed + 100)
    h = ctx.ransac_hypotheses(fp["p1"], fp["p2"], tent, 8, iters, thr, seed + 100)
    assert np.array_equal(h["sets"], o["sets"])
    assert np.array_equal(bits(h["F_all"]), bits(o["F_all"]))
    assert np.array_equal(h["cnt_all"], o["cnt_all"])
    assert np.array_equal(bits(h["score_all"]), bits(o["score_all"]))
    assert g["rc"] == 0 and g["best"] == o["best"]
    assert np.array_equal(bits(g["F"]), bits(o["F"]))
    assert rel_frob(g["F"], o["F"]) <= 1e-5
    assert np.array_equal(g["mask"], o["mask"])
    assert g["n_inliers"] == o["n_inliers"] and bits(np.array([g["score"]]))[0] == bits(np.array([o["score"]]))[0]


def test_find_fundamental_tie_rule(ctx, oracle):
    fp = synth.frame_pair(300, 9, noise_px=0.0, outlier_frac=0.0)
    tent = np.stack([np.arange(300), fp["gt"]], 1).astype(np.int32)
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, 256, 10.0, 77, want_all=True)
    assert (o["cnt_all"] == o["cnt_all"].max()).sum() > 1
    g = ctx.ransac_fundamental(fp["p1"], fp["p2"], tent, 8, 256, 10.0, 77)
    assert g["best"] == o["best"] and np.array_equal(g["mask"], o["mask"]) and np.array_equal(bits(g["F"]), bits(o["F"]))


def test_find_fundamental_edge_cases(ctx, oracle):
    fp = synth.frame_pair(100, 1)
    tent = oracle.match_hamming(fp["d1"], fp["d2"])
    # fewer matches than min_items: reference is UB; ABI reports TOO_FEW and writes nothing
    g = ctx.ransac_fundamental(fp["p1"], fp["p2"], tent[:5], 8, 10, 10.0, 1)
    assert g["rc"] == 3 and g["best"] == -1 and not g["F"].any()
    # exactly min_items matches
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent[:8], 8, 16, 10.0, 4)
    g = ctx.ransac_fundamental(fp["p1"], fp["p2"], tent[:8], 8, 16, 10.0, 4)
    assert g["best"] == o["best"] and np.array_equal(g["mask"], o["mask"])
    # min_items < 8: trailing sample slots stay 0 (src/RansacFilter.cpp:17)
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 5, 32, 10.0, 6, want_all=True)
    h = ctx.ransac_hypotheses(fp["p1"], fp["p2"], tent, 5, 32, 10.0, 6)
    assert np.array_equal(h["sets"], o["sets"]) and (h["sets"][:, 5:] == 0).all()
    assert np.array_equal(bits(h["F_all"]), bits(o["F_all"]))
    # threshold so small nothing is an inlier: winner decided by the score-only rule from (0, 0)
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, 32, 0.0, 2)
    g = ctx.ransac_fundamental(fp["p1"], fp["p2"], tent, 8, 32, 0.0, 2)
    assert g["best"] == o["best"] and g["n_inliers"] == o["n_inliers"] and np.array_equal(g["mask"], o["mask"])
    # all points identical: every residual is 0/0 = NaN, nothing is ever accepted
    z1 = np.full((20, 2), 5.0, np.float32)
    mm = np.stack([np.arange(20), np.arange(20)], 1).astype(np.int32)
    o = oracle.find_fundamental(z1, z1, mm, 8, 16, 10.0, 2)
    g = ctx.ransac_fundamental(z1, z1, mm, 8, 16, 10.0, 2)
    assert (g["rc"] == 5) == (o["best"] < 0) and g["best"] == o["best"]


# ---------------------------------------------------------------- Hamming matcher
def test_knn2_hamming_golden(ctx, golden):
    for name in ("knn", "knnc"):
        idx, dist = ctx.knn2_hamming(golden[f"{name}_d1"], golden[f"{name}_d2"])
        assert np.array_equal(idx, golden[f"{name}_idx"])                      # cv2 BFMatcher order
        assert np.array_equal(dist.astype(np.float32), golden[f"{name}_dist"])
        pairs = ctx.match_hamming(golden[f"{name}_d1"], golden[f"{name}_d2"], 0.7)
        q = np.nonzero(golden[f"{name}_keep"])[0]
        assert np.array_equal(pairs[:, 0], q) and np.array_equal(pairs[:, 1], golden[f"{name}_idx"][q, 0])


@pytest.mark.parametrize("n1,n2,nbytes", [(1, 2, 32), (3, 5, 32), (255, 257, 32), (1000, 129, 32), (2000, 2000, 32),
                                          (5000, 5000, 32), (700, 900, 16), (300, 400, 64)])
def test_knn2_hamming_vs_oracle(ctx, oracle, n1, n2, nbytes):
    rng = np.random.default_rng(n1 * 7 + n2)
    d2 = synth.random_descriptors(rng, n2, nbytes)
    d1 = synth.random_descriptors(rng, n1, nbytes)
    k = min(n1, n2) // 2
    d1[:k] = synth.flip_bits(rng, d2[rng.permutation(n2)[:k]], 12)
    if n2 > 40:
        d2[n2 - 20:] = d2[:20]         # duplicated train rows: equal distances, lower index must win
    idx, dist = ctx.knn2_hamming(d1, d2)
    oidx, odist = oracle.knn2_hamming(d1, d2)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    for ratio in (0.7, 0.9, 1.0):
        assert np.array_equal(ctx.match_hamming(d1, d2, ratio), oracle.match_hamming(d1, d2, ratio))


TC_SHAPES = [(1, 2), (1, 7), (3, 8), (5, 9), (255, 239), (256, 240), (257, 241), (1000, 129), (1000, 479), (255, 481),
             (256, 8), (300, 1000), (3000, 239), (3000, 241), (1000, 16383), (2500, 16384), (513, 720), (100, 961)]


@pytest.mark.parametrize("drain", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
@pytest.mark.parametrize("n1,n2", TC_SHAPES)
def test_knn2_hamming_tensor_path_edge_shapes(ctx, oracle, n1, n2, drain):
    """The tcgen05 matcher (k_knn2_tc4 + k_knn2_tc_fix) forced onto shapes the size heuristic would send to the popcount
    kernel: n1 != n2, n2 around the 240-column tile and the 8-column group, n2 < one tile, n1 around the 256-query block, the
    n2 = 16 384 key limit; duplicated train rows straddling group and tile boundaries (ties to the lower index).
    BFMatcher knnMatch k = 2 order, reference src/Frame.cpp:83-85; ratio test :91."""
    ctx.set_option("hamming_tc", 1)
    ctx.set_option("tc_drain", drain)
    rng = np.random.default_rng(n1 * 131 + n2)
    d2 = synth.random_descriptors(rng, n2)
    d1 = synth.random_descriptors(rng, n1)
    k = min(n1, n2) // 2
    if k:
        d1[:k] = synth.flip_bits(rng, d2[rng.permutation(n2)[:k]], 12)
    # equal rows across a group boundary (7|8), a tile boundary (239|240) and the last valid column
    for a, b in ((7, 8), (0, 15), (239, 240), (232, 479), (n2 - 1, 3), (n2 - 2, n2 - 9)):
        if 0 <= a < n2 and 0 <= b < n2 and a != b:
            d2[max(a, b)] = d2[min(a, b)]
    if n1 > 4 and n2 > 16:
        d1[1] = d2[7]      # zero distance to a duplicated pair
        d1[2] = d2[n2 - 1]
    idx, dist = ctx.knn2_hamming(d1, d2)
    oidx, odist = oracle.knn2_hamming(d1, d2)
    assert np.array_equal(dist, odist)
    assert np.array_equal(idx, oidx)
    for ratio in (0.7, 1.0):
        assert np.array_equal(ctx.match_hamming(d1, d2, ratio), oracle.match_hamming(d1, d2, ratio))


@pytest.mark.parametrize("drain", [6, 7, 8, 9, 10, 106, 109])   # 100 + d: variant d with two UMMA-issuing warps (tc_issuers = 2)
@pytest.mark.parametrize("n2", [500, 4999, 6001, 7680, 7681])
def test_packed_drain_interleaved_group_ties(ctx, oracle, n2, drain):
    """The packed drain (tc_drain = 6, the default) ranks groups of 8 same-parity columns of a 16-column span; the even and the
    odd group of one span interleave, so equal best distances in both must still resolve to the lower train index, for the best
    and for the second neighbour's distance. Every query here has its nearest and/or second-nearest distance duplicated across
    such a pair of groups (and across spans, parts and tiles); vb_match_hamming at ratios that make the second distance
    decide. n2 = 7 681 is one tile over the packed key's range and takes the float drain (6 001: over variant 7's, takes 6). Reference src/Frame.cpp:83-95."""
    ctx.set_option("hamming_tc", 1)
    ctx.set_option("tc_issuers", 2 if drain >= 100 else 1)
    drain %= 100
    ctx.set_option("tc_drain", drain)
    rng = np.random.default_rng(n2)
    n1 = 700
    d2 = synth.random_descriptors(rng, n2)
    d1 = synth.random_descriptors(rng, n1)
    # column pairs: siblings of one span (odd below even, even below odd), neighbours across a span / part / tile boundary
    pairs = [(1, 4), (2, 3), (17, 30), (15, 16), (63, 64), (239, 240), (33, 36), (250, 265), (n2 - 3, n2 - 2), (n2 - 16, n2 - 1)]
    for i, (a, b) in enumerate(pairs):
        base = synth.random_descriptors(rng, 1)[0]
        for rep in range(8):
            q = 10 * i + rep
            a_, b_ = (a + 16 * rep * 3) % n2, (b + 16 * rep * 3) % n2
            if a_ == b_:
                continue
            d2[a_] = base
            d2[b_] = base
            if rep % 2 == 0:      # nearest distance duplicated (nearest = lower index, second distance = the same value)
                d1[q] = synth.flip_bits(rng, base[None], 6)[0]
            else:                 # a unique nearest elsewhere, the duplicated pair ties for second place
                near = synth.flip_bits(rng, base[None], 3)[0]
                d1[q] = near
                d2[(a_ + 97) % n2] = near
    for ratio in (0.5, 0.7, 0.99, 1.0):
        assert np.array_equal(ctx.match_hamming(d1, d2, ratio), oracle.match_hamming(d1, d2, ratio)), ratio
    ctx.set_option("tc_drain", 1)
    want = ctx.match_hamming(d1, d2, 0.8)
    ctx.set_option("tc_drain", drain)
    assert np.array_equal(ctx.match_hamming(d1, d2, 0.8), want)


@pytest.mark.parametrize("k", [241, 600, 1000])
def test_tensor_path_whole_pair_and_sequence_edge_shapes(ctx, oracle, k):
    """match_features (P = 1) and a 4-frame sequence (P = 3, shared expanded frames) through the forced tensor path at sizes
    that leave a ragged last tile / query block."""
    ctx.set_option("hamming_tc", 1)
    pts, desc = synth.sequence(4, k, k)
    prm = ctx.params(0.7, 8, 64, 10.0, 77)
    res, out = ctx.pairs_run(pts, desc, prm)
    for i in range(3):
        o = oracle.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, 64, 10.0, 77 + i)
        assert res["n_tentative"][i] == o["n_tentative"] and res["n_matches"][i] == max(o["n"], 0)
        if o["n"] > 0:
            assert np.array_equal(out[i, :o["n"]], o["matches"]) and np.array_equal(bits(res["F"][i]), bits(o["F"].reshape(-1)))
    g = ctx.match_features(pts[0], desc[0], pts[2][:k - 37], desc[2][:k - 37], prm)     # n1 != n2
    o = oracle.match_features(pts[0], desc[0], pts[2][:k - 37], desc[2][:k - 37], 0.7, 8, 64, 10.0, 77)
    assert g["n_tentative"] == o["n_tentative"] and g["n"] == max(o["n"], 0)
    if o["n"] > 0:
        assert np.array_equal(g["matches"], o["matches"]) and np.array_equal(bits(g["F"]), bits(o["F"]))


def test_hamming_too_few_train(ctx):
    from vslam_b200.lib import VbError
    d = np.zeros((4, 32), np.uint8)
    with pytest.raises(VbError) as e:
        ctx.knn2_hamming(d, d[:1])
    assert e.value.code == 3


# ---------------------------------------------------------------- whole pair / sequence
@pytest.mark.parametrize("k,iters,seed", [(2000, 100, 0), (2000, 1024, 3), (5000, 1024, 1)])
def test_match_features_bit_exact(ctx, oracle, k, iters, seed):
    fp = synth.frame_pair(k, seed)
    prm = ctx.params(0.7, 8, iters, 10.0, 1234 + seed)
    g = ctx.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], prm)
    o = oracle.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], 0.7, 8, iters, 10.0, 1234 + seed)
    assert g["status"] == 0 and g["n_tentative"] == o["n_tentative"] and g["best"] == o["best"]
    assert g["n"] == o["n"] and np.array_equal(g["matches"], o["matches"])
    assert np.array_equal(bits(g["F"]), bits(o["F"])) and rel_frob(g["F"], o["F"]) <= 1e-5
    assert o["n"] > 0.4 * k


def test_pairs_run_sequence_bit_exact(ctx, oracle):
    nframes, k = 9, 1500
    pts, desc = synth.sequence(nframes, k, 42)
    prm = ctx.params(0.7, 8, 256, 10.0, 500)
    res, out = ctx.pairs_run(pts, desc, prm)
    for i in range(nframes - 1):
        o = oracle.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, 256, 10.0, 500 + i)
        assert res["status"][i] == 0
        assert res["n_tentative"][i] == o["n_tentative"] and res["best_hyp"][i] == o["best"]
        assert res["n_matches"][i] == o["n"]
        assert np.array_equal(out[i, :o["n"]], o["matches"])
        assert np.array_equal(bits(res["F"][i]), bits(o["F"].reshape(-1)))
    # no-download variant reports the same counts
    res2, _ = ctx.pairs_run(pts, desc, prm, want_matches=False)
    assert np.array_equal(res2["n_inliers"], res["n_inliers"]) and np.array_equal(res2["best_hyp"], res["best_hyp"])


def test_pairs_run_degenerate_pair_in_batch(ctx, oracle):
    """A pair whose descriptors do not match at all (< 8 tentative matches) must not disturb its neighbours."""
    nframes, k = 5, 600
    pts, desc = synth.sequence(nframes, k, 7)
    desc[2] = synth.random_descriptors(np.random.default_rng(1), k)      # frame 2 unrelated to 1 and 3
    prm = ctx.params(0.7, 8, 64, 10.0, 9)
    res, out = ctx.pairs_run(pts, desc, prm)
    for i in range(nframes - 1):
        o = oracle.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, 64, 10.0, 9 + i)
        if o["n"] < 0:
            assert res["status"][i] == 3 and res["n_matches"][i] == 0
        else:
            assert res["status"][i] == 0 and res["n_matches"][i] == o["n"] and np.array_equal(out[i, :o["n"]], o["matches"])
    assert (res["status"] == 3).any()


# ---------------------------------------------------------------- KD-tree
@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 1023, 1024, 1025, 5000, 6144, 6145, 7000, 12289, 20000, 100000])
def test_kdtree_build_bit_exact(ctx, oracle, n):
    pts = synth.frame_pair(max(n, 8), n + 1)["p1"][:n].copy()
    t = ctx.kdtree_build(pts)
    idx, pre = t.export()
    opre = oracle.kdtree_build(pts)
    assert np.array_equal(idx.astype(np.int32), opre)
    assert np.array_equal(pre, pts[opre])
    assert t.height == oracle.lib.vbo_kdtree_height(n)
    t.free()


@pytest.mark.parametrize("n", [2700, 9000, 30000])
def test_kdtree_build_with_ties(ctx, oracle, n):
    """Integer grid coordinates as in the reference's own test (tests/test_kdtree.cpp:39-45): massive ties. The documented
    tie order (coordinate, then original index) must match the oracle's — in the one-CTA build and through the top-down
    levels of a large tree (radix-selected medians inside runs of equal coordinates, stable partitions)."""
    rng = np.random.default_rng(n)
    pts = rng.integers(0, 100, (n, 2)).astype(np.float32)
    pts[: n // 10] = pts[0]                                   # and one long run of identical points
    t = ctx.kdtree_build(pts)
    idx, _ = t.export()
    assert np.array_equal(idx.astype(np.int32), oracle.kdtree_build(pts))
    assert sorted(idx.tolist()) == list(range(n))


@pytest.mark.parametrize("n", [1, 5, 3000, 5000, 20000])
def test_kdtree_queries_bit_exact(ctx, oracle, n):
    rng = np.random.default_rng(n)
    pts = synth.frame_pair(max(n, 8), n + 3)["p1"][:n].copy()
    pre = oracle.kdtree_build(pts)
    t = ctx.kdtree_build(pts)
    nq = 600
    q = np.ascontiguousarray(pts[rng.integers(0, n, nq)] + rng.uniform(-3, 3, (nq, 2)), np.float32)
    q[:20] = pts[rng.integers(0, n, 20)]                       # exact hits
    pt, idx, d2 = t.nearest(q)
    for i in range(nq):
        slot, od2 = oracle.kdtree_nearest(pts, pre, q[i])
        assert idx[i] == pre[slot] and d2[i] == od2 and np.array_equal(pt[i], pts[pre[slot]])
    pt0, idx0, _ = t.nearest(q[:50], 1e-12)                    # bounded search: mostly nothing in range
    for i in range(50):
        slot, _ = oracle.kdtree_nearest(pts, pre, q[i], 1e-12)
        assert idx0[i] == (pre[slot] if slot >= 0 else -1)
        if slot < 0:
            assert np.array_equal(pt0[i], [0, 0])              # src/KDTree.cpp:38-42
    for r in (2.0, 17.5, 60.0):
        off, out = t.radius(q[:200], r)
        for i in range(200):
            oidx, c = oracle.kdtree_radius(pts, pre, q[i], r)
            assert off[i + 1] - off[i] == c
            assert np.array_equal(out[off[i]:off[i + 1]].astype(np.int32), oidx)     # same order (pre-order)
    t.free()


@pytest.mark.parametrize("lanes", [8, 32])
def test_kdtree_nearest_lane_mappings_agree(ctx, oracle, lanes):
    """Option kd_lanes_per_query (the warp-per-query mapping of the north star, kept as an A/B): same results as the
    thread-per-query kernel and the oracle (src/KDTree.cpp:45-71 visiting order)."""
    rng = np.random.default_rng(lanes)
    n = 5000
    pts = synth.frame_pair(n, 77)["p1"]
    pre = oracle.kdtree_build(pts)
    t = ctx.kdtree_build(pts)
    q = np.ascontiguousarray(pts[rng.integers(0, n, 3000)] + rng.uniform(-3, 3, (3000, 2)), np.float32)
    pt1, idx1, d1 = t.nearest(q)
    ctx.set_option("kd_lanes_per_query", lanes)
    pt2, idx2, d2 = t.nearest(q)
    assert np.array_equal(idx1, idx2) and np.array_equal(bits(d1), bits(d2)) and np.array_equal(bits(pt1), bits(pt2))
    for i in range(0, 3000, 60):
        slot, od2 = oracle.kdtree_nearest(pts, pre, q[i])
        assert idx2[i] == pre[slot] and d2[i] == od2
    t.free()


@pytest.mark.parametrize("n,k", [(1, 1), (5, 8), (3000, 1), (3000, 2), (5000, 8), (20000, 32)])
def test_kdtree_knn_bit_exact(ctx, oracle, n, k):
    """k nearest neighbours (build-defined generalisation of `nearest`, src/KDTree.cpp:45-71; the reference's k_nearest is a
    commented-out declaration): indices, distances and counts equal the oracle's; k = 1 equals vb_kdtree_nearest; against
    brute force the distance lists agree; a finite max_d2 truncates the list."""
    rng = np.random.default_rng(n + k)
    pts = synth.frame_pair(max(n, 8), n + 5)["p1"][:n].copy()
    pre = oracle.kdtree_build(pts)
    t = ctx.kdtree_build(pts)
    nq = 400
    q = np.ascontiguousarray(pts[rng.integers(0, n, nq)] + rng.uniform(-4, 4, (nq, 2)), np.float32)
    q[:10] = pts[rng.integers(0, n, 10)]
    idx, d2, cnt = t.knn(q, k)
    for i in range(nq):
        slot, od2 = oracle.kdtree_knn(pts, pre, q[i], k)
        assert cnt[i] == len(slot) == min(k, n)
        assert np.array_equal(idx[i, :cnt[i]], pre[slot]) and np.array_equal(bits(d2[i, :cnt[i]]), bits(od2))
        assert (idx[i, cnt[i]:] == -1).all()
    dd = ((pts[None, :, :].astype(np.float32) - q[:50, None, :]) ** 2)
    bf = np.sort((dd[..., 0] + dd[..., 1]).astype(np.float32), axis=1)[:, :min(k, n)]
    assert np.allclose(d2[:50, :min(k, n)], bf, rtol=1e-6)
    if k == 1:
        _, i1, e1 = t.nearest(q)
        assert np.array_equal(i1, idx[:, 0]) and np.array_equal(bits(e1), bits(d2[:, 0]))
    idx2, d22, cnt2 = t.knn(q, k, 9.0)                                  # bounded: only points within 3 px
    for i in range(0, nq, 7):
        slot, od2 = oracle.kdtree_knn(pts, pre, q[i], k, 9.0)
        assert cnt2[i] == len(slot) and np.array_equal(idx2[i, :cnt2[i]], pre[slot]) and (d22[i, :cnt2[i]] < 9.0).all()
    t.free()


def test_kdtree_reference_test_protocol(ctx):
    """The reference's own acceptance test (tests/test_kdtree.cpp:47-146) replayed against the GPU tree:
    integer points in [0,100)^2, nearest accepted on equal distance, radius results compared as sets."""
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(2500, 3000))
        pts = rng.integers(0, 100, (n, 2)).astype(np.float32)
        t = ctx.kdtree_build(pts)
        q = rng.integers(0, 100, (16, 2)).astype(np.float32)
        pt, idx, d2 = t.nearest(q)
        dd = ((pts[None, :, :] - q[:, None, :]) ** 2).sum(-1)
        assert np.array_equal(d2, dd.min(1).astype(np.float32))
        r = float(rng.uniform(10, 100))
        off, out = t.radius(q, r)
        for i in range(len(q)):
            assert sorted(out[off[i]:off[i + 1]].tolist()) == np.nonzero(dd[i] < np.float32(r) * np.float32(r))[0].tolist()
        t.free()


def test_kdtree_radius_capacity_contract(ctx):
    pts = synth.frame_pair(1000, 1)["p1"]
    t = ctx.kdtree_build(pts)
    q = np.ascontiguousarray(pts[:10])
    off, out, tot = np.zeros(11, np.uint32), np.zeros(4, np.uint32), C.c_uint64()
    rc = ctx.L.vb_kdtree_radius(t.h, q.ctypes.data_as(C.c_void_p), 10, 500.0, off.ctypes.data_as(C.c_void_p),
                                out.ctypes.data_as(C.c_void_p), 4, C.byref(tot))
    assert rc == 4 and tot.value > 4 and off[10] == tot.value
    t.free()


# ---------------------------------------------------------------- float descriptors (config 3)
@pytest.mark.parametrize("n1,n2,dim", [(300, 500, 128), (1000, 1000, 128), (200, 333, 64)])
def test_knn2_l2f_bit_exact(ctx, oracle, n1, n2, dim):
    fp = synth.frame_pair_float(max(n1, n2), 4, dim=dim)
    d1, d2 = np.ascontiguousarray(fp["d1"][:n1]), np.ascontiguousarray(fp["d2"][:n2])
    d2[-5:] = d2[:5]
    idx, dist = ctx.knn2_l2f(d1, d2)
    oidx, odist = oracle.knn2_l2f(d1, d2)
    assert np.array_equal(idx, oidx) and np.array_equal(bits(dist), bits(odist))
    assert np.array_equal(ctx.match_l2f(d1, d2, 0.7), oracle.match_l2f(d1, d2, 0.7))


def _float_descriptors(kind, n1, n2, dim, seed):
    rng = np.random.default_rng(seed)
    d1 = rng.standard_normal((n1, dim)).astype(np.float32)
    d2 = rng.standard_normal((n2, dim)).astype(np.float32)
    if kind == "unit":            # unit-norm, half of the queries are noisy copies of train rows
        d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
        d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        m = min(n1, n2) // 2
        d1[:m] = d2[:m] + 0.05 * rng.standard_normal((m, dim)).astype(np.float32)
    elif kind == "sift":          # integer-valued, large norms: many exactly equal distances
        d1, d2 = np.abs(d1 * 40).round().astype(np.float32), np.abs(d2 * 40).round().astype(np.float32)
    elif kind == "near":          # train rows = a few prototypes + perturbations far below bf16 resolution: the GEMM cannot order them
        proto = rng.standard_normal((7, dim)).astype(np.float32)
        d2 = proto[rng.integers(0, 7, n2)] + (1e-5 * rng.standard_normal((n2, dim))).astype(np.float32)
        d1 = proto[rng.integers(0, 7, n1)] + (1e-5 * rng.standard_normal((n1, dim))).astype(np.float32)
    elif kind == "equal":         # every train row equal but one: every chunk minimum ties, exact scan fallback
        d2[:] = d2[0]
        d2[min(7, n2 - 1)] += 1e-3
    if n2 > 40:                   # duplicated train rows (tie order) and a zero distance
        d2[n2 // 2] = d2[3]
        d1[min(5, n1 - 1)] = d2[3]
    return np.ascontiguousarray(d1), np.ascontiguousarray(d2)


@pytest.mark.parametrize("kind,n1,n2,dim", [("unit", 256, 256, 128), ("unit", 300, 700, 64), ("gauss", 1000, 513, 128),
                                            ("sift", 1500, 2100, 128), ("unit", 37, 2, 128), ("equal", 500, 40, 64),
                                            ("equal", 300, 1000, 128), ("sift", 700, 300, 64), ("near", 600, 900, 128),
                                            ("near", 257, 300, 64)])
def test_knn2_l2f_tensor_path_bit_exact(ctx, oracle, kind, n1, n2, dim):
    """The tcgen05 (bf16 GEMM + exact re-evaluation) path returns the oracle's indices and distance bits."""
    ctx.set_option("l2_tc", 1)
    d1, d2 = _float_descriptors(kind, n1, n2, dim, 11)
    idx, dist = ctx.knn2_l2f(d1, d2)
    oidx, odist = oracle.knn2_l2f(d1, d2)
    assert np.array_equal(idx, oidx) and np.array_equal(bits(dist), bits(odist))
    assert np.array_equal(ctx.match_l2f(d1, d2, 0.7), oracle.match_l2f(d1, d2, 0.7))


def test_knn2_l2f_tensor_path_equals_exact_kernel_at_size(ctx, oracle):
    """6000 x 7000 x 128: tensor path == exact SIMT kernel on every query, == oracle on a slice."""
    d1, d2 = _float_descriptors("unit", 6000, 7000, 128, 3)
    ctx.set_option("l2_tc", 1)
    it, dt = ctx.knn2_l2f(d1, d2)
    ctx.set_option("l2_tc", 0)
    ie, de = ctx.knn2_l2f(d1, d2)
    assert np.array_equal(it, ie) and np.array_equal(bits(dt), bits(de))
    oi, od = oracle.knn2_l2f(np.ascontiguousarray(d1[:200]), d2)
    assert np.array_equal(it[:200], oi) and np.array_equal(bits(dt[:200]), bits(od))


@pytest.mark.parametrize("k,iters,seed", [(600, 64, 1), (3000, 256, 2), (6000, 512, 3)])
def test_match_features_l2f_whole_pair(ctx, oracle, k, iters, seed):
    """Float-descriptor match_features in ONE call (matcher -> ratio -> find_fundamental -> inlier copy-out on the device)
    == oracle matcher + oracle find_fundamental (src/Frame.cpp:91-102, src/RansacFilter.cpp:36-67), bit for bit."""
    fp = synth.frame_pair_float(k, seed)
    prm = ctx.params(0.7, 8, iters, 10.0, 900 + seed)
    g = ctx.match_features_l2f(fp["p1"], fp["d1"], fp["p2"], fp["d2"], prm)
    tent = oracle.match_l2f(fp["d1"], fp["d2"], 0.7)
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, iters, 10.0, 900 + seed)
    assert g["status"] == 0 and g["n_tentative"] == len(tent) and g["best"] == o["best"]
    assert g["n_inliers"] == o["n_inliers"] and np.array_equal(bits(g["F"]), bits(o["F"]))
    assert np.array_equal(g["matches"], tent[o["mask"].astype(bool)])
    assert g["n"] > 0.3 * k


# ---------------------------------------------------------------- counting kernel (the pair pipeline's scorer)
@pytest.mark.parametrize("m,H,thr", [(1, 3, 10.0), (129, 70, 10.0), (3500, 1024, 10.0), (3500, 300, 0.2), (20000, 257, 10.0),
                                     (8193, 64, 1e-6), (3000, 64, 1e6)])
def test_counts_equal_exact_scoring(ctx, oracle, m, H, thr):
    """k_count (relaxed arithmetic + error bound + exact fallback) returns the reference's counts for every hypothesis."""
    corr = synth.correspondences(max(m, 8), m + H)[:m].copy()
    Fs = _hyp_bank(oracle, synth.correspondences(512, 3), H, m)
    cnt = ctx.ransac_counts(corr, Fs, thr)
    ecnt, _ = ctx.ransac_score(corr, Fs, thr)
    assert np.array_equal(cnt, ecnt)
    pick = np.r_[0:min(H, 6)]
    ocnt, _ = _oracle_scores(oracle, corr, Fs[pick], thr)
    assert np.array_equal(cnt[pick], ocnt)


def test_counts_threshold_exactly_on_residuals(ctx, oracle):
    """Adversarial thresholds: thr set to residual values that actually occur (e == thr must count, the next float below
    must not), so the decision can only come out right through the exact fallback."""
    corr = synth.correspondences(2000, 21)
    Fs = _hyp_bank(oracle, corr, 40, 5)
    p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
    mm = np.stack([np.arange(len(corr)), np.arange(len(corr))], 1).astype(np.int32)
    rng = np.random.default_rng(2)
    for h in (0, 7, 19, 33):
        _, e, _, _ = oracle.residual(p1, p2, mm, Fs[h].reshape(3, 3), 10.0)
        finite = e[np.isfinite(e) & (e > 0)]
        for thr in rng.choice(finite, 6, replace=False):
            for t in (np.float32(thr), np.nextafter(np.float32(thr), np.float32(0)), np.nextafter(np.float32(thr), np.float32(np.inf))):
                cnt = ctx.ransac_counts(corr, Fs, float(t))
                ecnt, _ = ctx.ransac_score(corr, Fs, float(t))
                assert np.array_equal(cnt, ecnt)
                assert cnt[h] == int((e <= t).sum())


def test_counts_degenerate_models_and_coordinates(ctx, oracle, golden):
    """Zero / NaN / inf / huge F entries and the golden set's x/0, 0/0 cases: everything uncertain goes to the exact path."""
    g = golden
    p1, p2, mm = g["res_p1"], g["res_p2"], g["res_matches"]
    corr = np.ascontiguousarray(np.concatenate([p1[mm[:, 0]], p2[mm[:, 1]]], 1), np.float32)
    Fs = g["res_F"].reshape(-1, 9).astype(np.float32)
    cnt = ctx.ransac_counts(corr, Fs, float(g["res_thr"]))
    assert np.array_equal(cnt, g["res_cnt"])
    corr2 = synth.correspondences(1500, 4)
    bank = _hyp_bank(oracle, corr2, 16, 1)
    weird = np.zeros((8, 9), np.float32)
    weird[1] = np.nan
    weird[2] = [0, 0, 0, 0, 0, 0, 0, 0, 1]            # a0 == 0 everywhere: x/0
    weird[3] = bank[0] * np.float32(1e30)
    weird[4] = bank[1] * np.float32(1e-30)
    weird[5] = [0, 0, 1e-38, 0, 0, 0, 0, 0, 1]        # a0^2 underflows
    weird[6] = np.inf
    weird[7] = bank[2]
    Fs2 = np.concatenate([bank, weird])
    for thr in (10.0, 0.0, np.inf):
        cnt = ctx.ransac_counts(corr2, Fs2, thr)
        ecnt, _ = ctx.ransac_score(corr2, Fs2, thr)
        assert np.array_equal(cnt, ecnt)
    big = corr2.copy()
    big[::7] *= np.float32(1e18)                      # squares overflow
    cnt = ctx.ransac_counts(big, Fs2, 10.0)
    ecnt, _ = ctx.ransac_score(big, Fs2, 10.0)
    assert np.array_equal(cnt, ecnt)


def test_pipeline_lazy_scores_equal_full_scoring(ctx, oracle):
    """vb_pairs_run with the counting kernel + tie scoring == the same call forced through full scoring, on data with many
    ties at the largest count (noise-free correspondences)."""
    pts, desc = synth.sequence(9, 1500, 3, noise_px=0.0, outlier_frac=0.2)
    prm = ctx.params(0.7, 8, 256, 10.0, 5)
    res_lazy, m_lazy = ctx.pairs_run(pts, desc, prm)
    ctx.set_option("ransac_lazy", 0)
    res_full, m_full = ctx.pairs_run(pts, desc, prm)
    for k in ("status", "n_tentative", "n_matches", "best_hyp", "n_inliers"):
        assert np.array_equal(res_lazy[k], res_full[k]), k
    assert np.array_equal(bits(res_lazy["score"]), bits(res_full["score"]))
    assert np.array_equal(bits(res_lazy["F"]), bits(res_full["F"]))
    for i, n in enumerate(res_lazy["n_matches"]):
        assert np.array_equal(m_lazy[i, :n], m_full[i, :n])
    assert (res_lazy["n_matches"] > 100).all()


def test_kdtree_batched_build_equals_single_builds(ctx, oracle):
    """vb_kdtree_build_batch_d: one tree per frame in one launch == the per-frame builds == the oracle (layout, radius order)."""
    import torch
    from vslam_b200.lib import KDTreeHandle
    nt, n = 6, 3000
    pts, _ = synth.sequence(nt, n, 12)
    pts_d = torch.from_numpy(pts).cuda()
    handles = (C.c_void_p * nt)()
    ctx._chk(ctx.L.vb_kdtree_build_batch_d(ctx.h, C.c_void_p(pts_d.data_ptr()), nt, n, handles))
    try:
        for i in range(nt):
            view = KDTreeHandle(ctx, None, n)      # non-owning wrapper around the batch's handle
            view.h = C.c_void_p(handles[i])
            idx, pp = view.export()
            assert np.array_equal(idx.astype(np.int32), oracle.kdtree_build(pts[i]))
            assert view.height == int(np.floor(np.log2(n))) + 1
            q = np.ascontiguousarray(pts[i][:50] + 0.3, np.float32)
            off, hits = view.radius(q, 3.0)
            pre = oracle.kdtree_build(pts[i])
            for j in range(50):
                oi, _ = oracle.kdtree_radius(pts[i], pre, q[j], 3.0)
                assert np.array_equal(hits[off[j]:off[j + 1]].astype(np.int32), oi)
            view.h = None                          # released with the batch below
    finally:
        ctx.L.vb_kdtree_free_batch(handles, nt)


@pytest.mark.parametrize("opts", [dict(count_packed=0, score_packed=0, hamming_fp4=0), dict(tc_fix8=0), dict(hamming_qpt=1, hamming_tc=0),
                                  dict(hamming_qpt=4, hamming_tc=0), dict(tc_drain=1), dict(tc_drain=2), dict(tc_drain=3),
                                  dict(tc_drain=4), dict(tc_drain=5), dict(tc_drain=6), dict(tc_drain=7), dict(tc_drain=8), dict(tc_drain=9),
                                  dict(tc_drain=10), dict(tc_drain=6, tc_issuers=2), dict(tc_fix_skip=0)])
def test_alternative_kernels_agree_with_oracle(ctx, oracle, opts):
    """The code paths behind vb_set_option — k_count<2> / k_score<2> (scalar-instruction versions), the fp8 matcher, the
    two-group fix pass on the match path, the popcount matcher's queries-per-thread variants — still agree with the oracle."""
    for name, v in opts.items():
        ctx.set_option(name, v)
    corr = synth.correspondences(3000, 5)
    rng = np.random.default_rng(1)
    Fs = np.stack([oracle.compute_fundamental(corr[s_, :2], corr[s_, 2:]).reshape(-1)
                   for s_ in (rng.choice(3000, 8, replace=False) for _ in range(300))])
    cnt, sc = ctx.ransac_score(corr, Fs, 10.0)
    cnt2 = ctx.ransac_counts(corr, Fs, 10.0)
    p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
    mm = np.stack([np.arange(3000), np.arange(3000)], 1).astype(np.int32)
    for h in range(0, 300, 37):
        _, _, n, s = oracle.residual(p1, p2, mm, Fs[h].reshape(3, 3), 10.0)
        assert cnt[h] == n == cnt2[h] and np.float32(s).view(np.uint32) == sc[h:h + 1].view(np.uint32)[0]
    assert np.array_equal(cnt, cnt2)
    fp = synth.frame_pair(2500, 3)
    g = ctx.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], ctx.params(0.7, 8, 256, 10.0, 9))
    o = oracle.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], 0.7, 8, 256, 10.0, 9)
    assert g["n"] == o["n"] and np.array_equal(g["matches"], o["matches"]) and np.array_equal(bits(g["F"]), bits(o["F"]))


def test_options_contract(ctx):
    from vslam_b200.lib import VbError
    with pytest.raises(VbError) as e:
        ctx.set_option("no_such_option", 1)
    assert e.value.code == 1
    with pytest.raises(VbError):
        ctx.set_option("tc_dbg", 2)          # timing-only options exist in a -DVB_TUNING build only


# ---------------------------------------------------------------- BASELINE configs 3 and 5 at full size
def test_config3_full_size(ctx, oracle):
    """BASELINE configs[2] as stated: 20 000 x 20 000 x 128-d float descriptors, 4 096 hypotheses. The tensor-core matcher
    equals the exact SIMT kernel on every query (indices and distance bits) and the oracle on a 400-query sample; the
    one-call whole pair equals oracle find_fundamental on that tentative list (src/Frame.cpp:83-102, src/RansacFilter.cpp:36-67)."""
    n, H = 20000, 4096
    fp = synth.frame_pair_float(n, 5)
    ctx.set_option("l2_tc", 1)
    it, dt = ctx.knn2_l2f(fp["d1"], fp["d2"])
    ctx.set_option("l2_tc", 0)
    ie, de = ctx.knn2_l2f(fp["d1"], fp["d2"])
    ctx.reset_options()
    assert np.array_equal(it, ie) and np.array_equal(bits(dt), bits(de))
    qs = np.arange(0, n, 50)
    oi, od = oracle.knn2_l2f(np.ascontiguousarray(fp["d1"][qs]), fp["d2"])
    assert np.array_equal(it[qs], oi) and np.array_equal(bits(dt[qs]), bits(od))
    prm = ctx.params(0.7, 8, H, 10.0, 77)
    g = ctx.match_features_l2f(fp["p1"], fp["d1"], fp["p2"], fp["d2"], prm)
    tent = ctx.match_l2f(fp["d1"], fp["d2"], 0.7)
    keep = (dt[:, 0].astype(np.float64) < dt[:, 1].astype(np.float64) * 0.7)          # the ratio test restated on the verified kNN
    assert np.array_equal(tent[:, 0], np.nonzero(keep)[0]) and np.array_equal(tent[:, 1], it[keep, 0])
    o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, H, 10.0, 77)
    assert g["status"] == 0 and g["n_tentative"] == len(tent) and g["best"] == o["best"] and g["n_inliers"] == o["n_inliers"]
    assert np.array_equal(bits(g["F"]), bits(o["F"])) and np.array_equal(g["matches"], tent[o["mask"].astype(bool)])
    assert g["n"] > 8000


@pytest.mark.parametrize("m,H", [(1000, 256), (1000000, 256), (1000, 16384), (1000000, 16384)])
def test_config5_corners(ctx, oracle, m, H):
    """BASELINE configs[4] corners through vb_ransac_score (count + residual sum of every hypothesis,
    src/RansacFilter.cpp:105-140): the oracle on a sample of hypotheses (bit-exact count and score), duplicates of a
    hypothesis give identical outputs, counts are additive over a split of the matches."""
    corr = synth.correspondences(m, 5)
    bank = _hyp_bank(oracle, synth.correspondences(4000, 3), 256, 9)
    rng = np.random.default_rng(H)
    pick = rng.integers(0, 256, H)
    Fs = np.ascontiguousarray(bank[pick])
    cnt, sc = ctx.ransac_score(corr, Fs, 10.0)
    sample = sorted(set([0, 1, H // 2, H - 1] + rng.integers(0, H, 4 if m > 100000 else 40).tolist()))
    ocnt, osc = _oracle_scores(oracle, corr, Fs[sample], 10.0)
    assert np.array_equal(cnt[sample], ocnt) and np.array_equal(bits(sc[sample]), bits(osc))
    first = {}
    for h, b in enumerate(pick):                 # every copy of a bank entry must report the same pair
        if b in first:
            assert cnt[h] == cnt[first[b]] and bits(sc[h:h + 1])[0] == bits(sc[first[b]:first[b] + 1])[0]
        else:
            first[b] = h
    cut = m // 3
    c1 = ctx.ransac_counts(np.ascontiguousarray(corr[:cut]), Fs, 10.0)
    c2 = ctx.ransac_counts(np.ascontiguousarray(corr[cut:]), Fs, 10.0)
    assert np.array_equal(cnt, c1 + c2)


# ---------------------------------------------------------------- opt-in mode: Hartley normalisation + true Sampson distance
@pytest.mark.parametrize("flags", [1, 2, 3])
@pytest.mark.parametrize("k,iters,thr", [(400, 64, 1.0), (3000, 512, 3.84), (2000, 1024, 0.5)])
def test_optin_mode_bit_exact_vs_oracle_mode(ctx, oracle, flags, k, iters, thr):
    """vb_ransac_fundamental_ex (NOT reference behaviour) == the oracle's mode of the same flags: winner,
    count, score, mask and F bit for bit. flags = 0 stays the reference path (every other test in this file)."""
    fp = synth.frame_pair(k, 31 + flags)
    tent = oracle.match_hamming(fp["d1"], fp["d2"])
    o = oracle.find_fundamental_ex(fp["p1"], fp["p2"], tent, 8, iters, thr, 11, flags)
    g = ctx.ransac_fundamental(fp["p1"], fp["p2"], tent, 8, iters, thr, 11, flags=flags)
    assert g["rc"] == 0 and g["best"] == o["best"] and g["n_inliers"] == o["n_inliers"]
    assert np.array_equal(g["mask"], o["mask"]) and np.array_equal(bits(g["F"]), bits(o["F"]))
    assert bits(np.array([g["score"]]))[0] == bits(np.array([o["score"]]))[0]
    assert flags == 1 or g["n_inliers"] > 0.5 * len(tent)     # (with the reference residual a threshold of 1 keeps few matches)


def test_optin_mode_repairs_sideways_motion(ctx, oracle):
    """The motion the reference's criterion is known to fail on (sideways translation): the default path finds a
    minority of the true inliers, the opt-in mode most of them; and flags = 0 through the _ex entry IS the default path."""
    R, t = synth._rot(0.004, -0.006, 0.003), np.array([0.35, 0.02, 0.05])
    rng = np.random.default_rng(1)
    k = 2000
    p1 = np.stack([rng.uniform(0, 1280, k), rng.uniform(0, 720, k)], 1)
    p2, _, ok = synth._advance(rng, p1, rng.uniform(4, 12, k), R, t, 0.5, 0.3)
    p1, p2 = p1.astype(np.float32), p2.astype(np.float32)
    mm = np.stack([np.arange(k), np.arange(k)], 1).astype(np.int32)
    d = ctx.ransac_fundamental(p1, p2, mm, 8, 512, 10.0, 5)
    r = ctx.ransac_fundamental(p1, p2, mm, 8, 512, 1.0, 5, flags=3)
    assert (d["mask"].astype(bool) == ok).mean() < 0.7 < 0.85 < (r["mask"].astype(bool) == ok).mean()
    o = oracle.find_fundamental_ex(p1, p2, mm, 8, 512, 1.0, 5, 3)
    assert np.array_equal(r["mask"], o["mask"]) and np.array_equal(bits(r["F"]), bits(o["F"]))
    from vslam_b200.lib import _ptr
    F0, m0 = np.zeros((3, 3), np.float32), np.zeros(k, np.uint8)
    n, s, b = C.c_int32(), C.c_float(), C.c_int32()
    rc = ctx.L.vb_ransac_fundamental_ex(ctx.h, _ptr(p1), k, _ptr(p2), k, _ptr(mm), k, 8, 512, 10.0, 5, 0, _ptr(F0), _ptr(m0),
                                        C.byref(n), C.byref(s), C.byref(b))
    assert rc == 0 and b.value == d["best"] and np.array_equal(m0, d["mask"]) and np.array_equal(bits(F0), bits(d["F"]))
