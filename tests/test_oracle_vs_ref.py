"""Pin the C oracle against the REFERENCE'S OWN sources built into oracle/_ref (dev container).

libvbref.so = /root/reference/src/{KDTree,RansacFilter}.cpp compiled unmodified against the
tests/cvlite OpenCV stand-in. These tests skip where it was not built.
"""
import os
import subprocess

import numpy as np
import pytest

from vslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_own_test_kdtree_passes(ref):
    exe = os.path.join(ROOT, "oracle", "_ref", "test_kdtree")
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if "successes out of" in l]
    assert lines == ["1000 successes out of 1000 trials"] * 2     # tests/test_kdtree.cpp:91,145


@pytest.mark.parametrize("n,mi,iters,seed", [(8, 8, 50, 1), (9, 8, 64, 2), (100, 8, 100, 42), (3000, 8, 1024, 7),
                                             (1000000, 8, 4096, 123), (20, 5, 30, 9), (50, 1, 10, 3)])
def test_sample_sets_match_libstdcxx(oracle, ref, n, mi, iters, seed):
    assert np.array_equal(oracle.initialize_sets(n, mi, iters, seed), ref.initialize_sets(n, mi, iters, seed))


def test_sample_sets_cover_rejection_path(oracle, ref):
    # with n ~ 1e6 the Lemire rejection branch (uniform_int_dist.h:_S_nd) fires about 2e-4 per draw
    from oracle_lib import Oracle  # noqa: F401
    import ctypes as C
    n, iters, seed = 1000000, 4096, 123
    # count rejections by replaying the raw stream
    class MT(C.Structure):
        _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]
    g = MT()
    oracle.lib.vbo_mt_seed(C.byref(g), seed)
    oracle.lib.vbo_mt_next.restype = C.c_uint32
    rej = 0
    for i in range(iters):
        for j in range(8):
            rng_ = n - j
            while True:
                low = (oracle.lib.vbo_mt_next(C.byref(g)) * rng_) & 0xFFFFFFFF
                if low < rng_ and low < ((1 << 32) - rng_) % rng_:
                    rej += 1
                    continue
                break
    assert rej > 0
    assert np.array_equal(oracle.initialize_sets(n, 8, iters, seed), ref.initialize_sets(n, 8, iters, seed))


@pytest.mark.parametrize("k,iters,thr,seed", [(400, 100, 10.0, 5), (2000, 100, 10.0, 0), (2000, 1024, 10.0, 1),
                                              (300, 64, 0.2, 3)])
def test_find_fundamental_matches_reference(oracle, ref, k, iters, thr, seed):
    fp = synth.frame_pair(k, seed)
    tent = oracle.match_hamming(fp["d1"], fp["d2"])
    assert len(tent) >= 8
    a = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, iters, thr, seed + 100)
    b = ref.find_fundamental(fp["p1"], fp["p2"], tent, 8, iters, thr, seed + 100)
    assert b["accepted"] == (a["best"] >= 0)
    assert np.array_equal(a["F"].view(np.uint32), b["F"].view(np.uint32))
    assert np.array_equal(a["mask"], b["mask"])


def test_find_fundamental_tie_rule_on_clean_data(oracle, ref):
    # noise-free, outlier-free: many hypotheses tie at the max inlier count, so the
    # "larger residual sum wins" rule (src/RansacFilter.cpp:59) decides
    fp = synth.frame_pair(300, 9, noise_px=0.0, outlier_frac=0.0)
    tent = np.stack([np.arange(300), fp["gt"]], 1).astype(np.int32)
    a = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, 256, 10.0, 77, want_all=True)
    b = ref.find_fundamental(fp["p1"], fp["p2"], tent, 8, 256, 10.0, 77)
    assert (a["cnt_all"] == a["cnt_all"].max()).sum() > 1
    assert np.array_equal(a["F"].view(np.uint32), b["F"].view(np.uint32)) and np.array_equal(a["mask"], b["mask"])


def test_residual_and_solve_match_reference(oracle, ref):
    fp = synth.frame_pair(500, 4)
    tent = oracle.match_hamming(fp["d1"], fp["d2"])
    rng = np.random.default_rng(0)
    for _ in range(10):
        sel = rng.choice(len(tent), 8, replace=False)
        p1s, p2s = fp["p1"][tent[sel, 0]], fp["p2"][tent[sel, 1]]
        F = oracle.compute_fundamental(p1s, p2s)
        assert np.array_equal(F.view(np.uint32), ref.compute_fundamental(p1s, p2s).view(np.uint32))
        mask, e, n, s = oracle.residual(fp["p1"], fp["p2"], tent, F, 10.0)
        rmask, rn, rs = ref.residual(fp["p1"], fp["p2"], tent, F, 10.0)
        assert np.array_equal(mask, rmask) and n == rn and s.view(np.uint32) == rs.view(np.uint32)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 2000, 5000])
def test_kdtree_build_matches_reference(oracle, ref, n):
    pts = synth.frame_pair(max(n, 8), n)["p1"][:n].copy()
    pre = oracle.kdtree_build(pts)
    rp, h, links_ok = ref.kdtree_build(pts)
    assert links_ok, "reference node array is not the pre-order layout the build relies on"
    assert np.array_equal(pts[pre], rp)
    assert h == oracle.lib.vbo_kdtree_height(n)
    if n <= 2000:     # the reference's frame_kdtree build copies all points per comparator (slow)
        fpre, fh = ref.frame_kdtree_build(pts)
        assert np.array_equal(pre.astype(np.int64), fpre) and fh == h


def test_kdtree_queries_match_reference(oracle, ref):
    rng = np.random.default_rng(3)
    pts = synth.frame_pair(3000, 12)["p1"]
    pre = oracle.kdtree_build(pts)
    q = np.ascontiguousarray(pts[rng.choice(3000, 400)] + rng.uniform(-3, 3, (400, 2)), np.float32)
    rn = ref.kdtree_nearest(pts, q)
    for i in range(len(q)):
        slot, d2 = oracle.kdtree_nearest(pts, pre, q[i])
        assert np.array_equal(pts[pre[slot]], rn[i])
    # bounded search: nothing within max_d2 -> reference returns {0,0} (src/KDTree.cpp:38-42)
    rn0 = ref.kdtree_nearest(pts, q, 1e-12)
    for i in range(50):
        slot, _ = oracle.kdtree_nearest(pts, pre, q[i], 1e-12)
        assert slot == -1 and np.array_equal(rn0[i], [0, 0])
    for r in (2.0, 17.5, 100.0):
        off, out = ref.frame_kdtree_radius(pts, q[:100], r)
        offv, outv = ref.kdtree_radius(pts, q[:100], r)
        for i in range(100):
            idx, c = oracle.kdtree_radius(pts, pre, q[i], r)
            assert c == off[i + 1] - off[i]
            assert np.array_equal(idx.astype(np.int64), out[off[i]:off[i + 1]])       # same order (pre-order)
            assert np.array_equal(pts[idx], outv[offv[i]:offv[i + 1]])
            d = np.linalg.norm(pts.astype(np.float64) - q[i].astype(np.float64), axis=1)
            assert set(idx.tolist()) >= set(np.nonzero(d < r - 1e-3)[0].tolist())
