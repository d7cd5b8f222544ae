"""Pin the C oracle against the cv2-generated golden vectors (tests/golden/gen_golden.py)."""
import numpy as np


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_residual_bit_exact_vs_cv2(oracle, golden):
    g = golden
    p1, p2, m, thr = g["res_p1"], g["res_p2"], g["res_matches"], float(g["res_thr"])
    assert len(g["res_F"]) >= 30
    saw_nan = saw_inf = False
    for i, F in enumerate(g["res_F"]):
        mask, e, n, score = oracle.residual(p1, p2, m, F, thr)
        assert np.array_equal(_bits(e), _bits(g["res_e"][i])), f"residual bits differ for F #{i}"
        assert np.array_equal(mask, g["res_mask"][i])
        assert n == int(g["res_cnt"][i])
        saw_nan |= bool(np.isnan(e).any())
        saw_inf |= bool(np.isinf(e).any())
        # cv::sum's order is SIMD-dependent: tolerance only (the oracle defines the order)
        ref = g["res_score"][i]
        if np.isfinite(ref):
            assert abs(float(score) - float(ref)) <= 2e-7 * abs(float(ref))
        else:
            assert (np.isnan(ref) and np.isnan(score)) or ref == score
    assert saw_nan and saw_inf, "golden set must cover 0/0 and x/0"


def test_gemm3_rule(golden):
    # the fp32 ((a0*b0 + a1*b1) + a2*b2) rule used for U*diag(D)*Vt (src/RansacFilter.cpp:101)
    f = np.float32
    for a, b, c in zip(golden["g3_a"], golden["g3_b"], golden["g3_c"]):
        r = np.empty((3, 3), f)
        for i in range(3):
            for j in range(3):
                r[i, j] = f(f(f(a[i, 0] * b[0, j]) + f(a[i, 1] * b[1, j])) + f(a[i, 2] * b[2, j]))
        assert np.array_equal(_bits(r), _bits(c))


def _sign_norm_dist(a, b):
    a = a.reshape(-1).astype(np.float64)
    b = b.reshape(-1).astype(np.float64)
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


def test_eight_point_vs_cv2_svd_tolerance(oracle, golden):
    """PARITY UNPINNED step: cv::SVDecomp on the un-normalised fp32 system is only accurate to about
    eps32*cond(A); the oracle's fp64 null vector must be at least as good a null vector and agree
    with cv2 to that level."""
    g = golden
    for A, F0, F2, p1s, p2s in zip(g["fm_A"], g["fm_F0"], g["fm_F2"], g["fm_p1"], g["fm_p2"]):
        f = oracle.null_vector(A)
        Ad = A.astype(np.float64)
        assert abs(np.linalg.norm(f.astype(np.float64)) - 1) < 1e-6
        assert np.linalg.norm(Ad @ f.astype(np.float64)) <= np.linalg.norm(Ad @ F0.reshape(-1).astype(np.float64)) * 1.5 + 1e-4
        assert _sign_norm_dist(f, F0) < 5e-3
        F = oracle.compute_fundamental(p1s, p2s)
        assert _sign_norm_dist(F, F2) < 5e-3
        s = np.linalg.svd(F.astype(np.float64), compute_uv=False)
        assert s[2] < 1e-6 * s[0]          # rank 2 enforced (:98-101)


def test_svd3_reconstructs(oracle):
    rng = np.random.default_rng(5)
    for _ in range(50):
        F = (rng.standard_normal((3, 3)) * 10.0 ** rng.uniform(-6, 2, (3, 3))).astype(np.float32)
        U, D, Vt = oracle.svd3(F)
        assert D[0] >= D[1] >= D[2] >= 0
        R = U.astype(np.float64) @ np.diag(D.astype(np.float64)) @ Vt.astype(np.float64)
        assert np.linalg.norm(R - F) <= 1e-6 * max(np.linalg.norm(F), 1e-30)
        assert np.allclose(np.linalg.svd(F.astype(np.float64), compute_uv=False), D, rtol=2e-6, atol=1e-7 * D[0])
    U, D, Vt = oracle.svd3(np.zeros((3, 3), np.float32))
    assert np.all(np.isfinite(U)) and np.all(D == 0)


def test_knn2_hamming_vs_cv2(oracle, golden):
    for name in ("knn", "knnc"):
        idx, dist = oracle.knn2_hamming(golden[f"{name}_d1"], golden[f"{name}_d2"])
        assert np.array_equal(idx, golden[f"{name}_idx"])
        assert np.array_equal(dist.astype(np.float32), golden[f"{name}_dist"])
        keep = np.array([oracle.lib.vbo_ratio_keep(int(a), int(b), 0.7) for a, b in dist], np.uint8)
        assert np.array_equal(keep, golden[f"{name}_keep"])
        pairs = oracle.match_hamming(golden[f"{name}_d1"], golden[f"{name}_d2"], 0.7)
        q = np.nonzero(golden[f"{name}_keep"])[0]
        assert np.array_equal(pairs[:, 0], q) and np.array_equal(pairs[:, 1], golden[f"{name}_idx"][q, 0])
    assert (golden["knnc_dist"][:, 0] == golden["knnc_dist"][:, 1]).any(), "tie coverage"


def test_ratio_integer_form_exhaustive(oracle):
    """src/Frame.cpp:91 compares float distances through a double product; for Hamming distances
    0..256 that is exactly 10*d0 < 7*d1 (the form the GPU matcher may use)."""
    for d0 in range(257):
        for d1 in range(257):
            assert bool(oracle.lib.vbo_ratio_keep(d0, d1, 0.7)) == (10 * d0 < 7 * d1) == (float(d0) < float(d1) * 0.7)


def test_score_sum_blocked_order(oracle):
    rng = np.random.default_rng(11)
    for n in (0, 1, 127, 128, 129, 8191, 8192, 8193, 20000):
        e = (10.0 ** rng.uniform(-6, 9, n)).astype(np.float32)
        tot = 0.0
        for g0 in range(0, n, 128 * 64):
            gs = 0.0
            for c0 in range(g0, min(g0 + 128 * 64, n), 128):
                cs = 0.0
                for v in e[c0:min(c0 + 128, g0 + 128 * 64, n)]:
                    cs += float(v)
                gs += cs
            tot += gs
        assert oracle.score_sum(e) == tot
        if n:
            assert abs(tot - float(np.sum(e.astype(np.float64)))) <= 1e-12 * tot


def test_projection_bit_exact_vs_cv2(oracle):
    """points * c2.t() (src/vslam.cpp:131) on both sides of OpenCV's 100-row small-matrix threshold, and K * R_t."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "projection_cv2_4_13.npz"))
    f = np.float32
    for tag in ("n1", "n7", "n99", "n100", "n101", "n640", "n5000"):
        X, c2, P = g[f"{tag}_X"], g[f"{tag}_c2"], g[f"{tag}_P"]
        assert np.array_equal(_bits(oracle.project_points(X, c2)), _bits(P)), tag
        # c2 = K * R_t.rowRange(0, 3): 3x3 times 3x4 without flags takes the fp32 small-matrix rule of test_gemm3_rule
        K, Rt = g["K"], g[f"{tag}_Rt"]
        r = np.empty((3, 4), f)
        for i in range(3):
            for j in range(4):
                r[i, j] = f(f(f(K[i, 0] * Rt[0, j]) + f(K[i, 1] * Rt[1, j])) + f(K[i, 2] * Rt[2, j]))
        assert np.array_equal(_bits(r), _bits(c2)), tag


def test_optin_mode_is_off_by_default_and_repairs_sideways_motion(oracle):
    """flags = 0 through vbo_find_fundamental_ex IS vbo_find_fundamental; Hartley + Sampson (flagged by the
    reference itself at src/RansacFilter.cpp:40 and :125-126) recovers the inliers of a sideways translation, which the
    reference's criterion cannot."""
    from vslam_b200 import synth
    R, t = synth._rot(0.004, -0.006, 0.003), np.array([0.35, 0.02, 0.05])
    rng = np.random.default_rng(1)
    k = 2000
    p1 = np.stack([rng.uniform(0, 1280, k), rng.uniform(0, 720, k)], 1)
    p2, _, ok = synth._advance(rng, p1, rng.uniform(4, 12, k), R, t, 0.5, 0.3)
    p1, p2 = p1.astype(np.float32), p2.astype(np.float32)
    mm = np.stack([np.arange(k), np.arange(k)], 1).astype(np.int32)
    a = oracle.find_fundamental(p1, p2, mm, 8, 256, 10.0, 5)
    b = oracle.find_fundamental_ex(p1, p2, mm, 8, 256, 10.0, 5, 0)
    assert a["best"] == b["best"] and np.array_equal(a["mask"], b["mask"]) and np.array_equal(a["F"].view(np.uint32), b["F"].view(np.uint32))
    r = oracle.find_fundamental_ex(p1, p2, mm, 8, 512, 1.0, 5, 3)
    assert (a["mask"].astype(bool) == ok).mean() < 0.7 and (r["mask"].astype(bool) == ok).mean() > 0.85
    Ft = synth.true_fundamental(R, t).reshape(-1)
    Fr = r["F"].reshape(-1).astype(np.float64)
    Fr /= np.linalg.norm(Fr)
    assert min(np.linalg.norm(Fr - Ft), np.linalg.norm(Fr + Ft)) < 0.05
    # the normalised solve of exact correspondences is the true F to fp32 accuracy where the raw-pixel solve is not
    good = np.nonzero(ok)[0][:8]
    X = np.stack([(p1[good, 0] - 640) / 525, (p1[good, 1] - 360) / 525, np.ones(8)], 1)
    Fh = oracle.compute_fundamental_hartley(p1[good], p2[good])
    assert abs(np.linalg.norm(Fh.astype(np.float64)) - 1.0) < 1e-6 and abs(np.linalg.det(Fh.astype(np.float64))) < 1e-8
    assert X.shape == (8, 3)
