"""The drop-in C++ boundary: include/KDTree.h + include/RansacFilter.h (same structs, class and free
functions as the reference's headers) driven by a C++ harness that uses them the way Frame.cpp /
vslam.cpp / tests/test_kdtree.cpp do, compared against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from vslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "adapter", "_build", "adapter_harness")


def build_harness():
    from vslam_b200 import lib as vl
    if not os.path.exists(vl.LIB_PATH):
        vl.build_library()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "adapter")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return HARNESS


def test_adapter_headers_compile_and_link():
    """CPU: the reference-shaped headers and the adapter sources compile (C++11, as the reference's makefile
    uses) and link against the C ABI."""
    assert os.path.exists(build_harness())


def test_adapter_headers_keep_reference_interface():
    """Every declaration the reference's callers rely on is present with the reference's spelling."""
    kd = open(os.path.join(ROOT, "include", "KDTree.h")).read()
    rf = open(os.path.join(ROOT, "include", "RansacFilter.h")).read()
    for needle in ["struct KDTree {", "struct frame_kdtree {", "cv::Point2f pt;", "usize pt_index;", "KDTreeNode *left;",
                   "KDTreeNode *right;", "KDTreeNode *root;", "u32 size = 0;", "u8 height = 0;",
                   "void construct_kdtree(KDTree &kdtree, const std::vector<cv::Point2f> &points);",
                   "cv::Point2f nearest(const KDTree &kdtree, const cv::Point2f &query_pt, float max_distance_sq = INFINITY);",
                   "std::vector<cv::Point2f> radius_search(const KDTree &kdtree, const cv::Point2f &query_pt, float radius);",
                   "void construct_kdtree(frame_kdtree &kdtree, const std::vector<cv::Point2f> &points);",
                   "std::vector<usize> radius_search(const frame_kdtree kdtree, const std::vector<cv::Point2f> &points,"]:
        assert needle in kd, needle
    for needle in ["class RansacFilter {", "const int min_items;", "const int max_iterations;", "const float threshold;",
                   "RansacFilter(const int min_items = 8, const int max_iterations = 100, const float threshold = 0.2);",
                   "void initialize_sets(const int n_matches);", "void find_fundamental(", "void compute_fundamental(",
                   "std::pair<int, float> compute_fundamental_residual("]:
        assert needle in rf, needle


def _run(tmp_path, p1, p2, matches, q, iters, min_items, seed, thr, radius):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<7i2f", len(p1), len(p2), len(matches), len(q), iters, min_items, seed, thr, radius))
        for a in (p1.astype(np.float32), p2.astype(np.float32), matches.astype(np.int32), q.astype(np.float32)):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([build_harness(), fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return open(fout, "rb").read()


class Reader:
    def __init__(self, b):
        self.b, self.o = b, 0

    def take(self, dtype, n):
        a = np.frombuffer(self.b, dtype, n, self.o)
        self.o += a.nbytes
        return a


@pytest.mark.gpu
@pytest.mark.parametrize("k,iters,seed", [(1500, 100, 3), (5000, 256, 11)])
def test_adapter_matches_oracle(tmp_path, oracle, k, iters, seed):
    fp = synth.frame_pair(k, seed)
    p1, p2 = fp["p1"], fp["p2"]
    tent = oracle.match_hamming(fp["d1"], fp["d2"])
    rng = np.random.default_rng(seed)
    nq = 64
    q = np.ascontiguousarray(p2[rng.integers(0, k, nq)] + rng.uniform(-2, 2, (nq, 2)), np.float32)
    radius, thr = 6.5, 10.0
    rd = Reader(_run(tmp_path, p1, p2, tent, q, iters, 8, 1000 + seed, thr, radius))
    pre = oracle.kdtree_build(p2)
    height = oracle.lib.vbo_kdtree_height(k)
    # value tree: node walk through left/right pointers == reference pre-order; size, height
    size, h = rd.take(np.int32, 2)
    assert (size, h) == (k, height)
    assert np.array_equal(rd.take(np.float32, 2 * k).reshape(k, 2), p2[pre])
    nn = rd.take(np.float32, 2 * nq).reshape(nq, 2)
    for i in range(nq):
        slot, _ = oracle.kdtree_nearest(p2, pre, q[i])
        assert np.array_equal(nn[i], p2[pre[slot]])
    for i in range(nq):
        c = int(rd.take(np.int32, 1)[0])
        pts = rd.take(np.float32, 2 * c).reshape(c, 2)
        oidx, oc = oracle.kdtree_radius(p2, pre, q[i], radius)
        assert c == oc and np.array_equal(pts, p2[oidx])
    # index tree
    size, h = rd.take(np.int32, 2)
    assert (size, h) == (k, height)
    assert np.array_equal(rd.take(np.int64, k), pre.astype(np.int64))
    for i in range(nq):
        c = int(rd.take(np.int32, 1)[0])
        idx = rd.take(np.int64, c)
        oidx, oc = oracle.kdtree_radius(p2, pre, q[i], radius)
        assert c == oc and np.array_equal(idx, oidx.astype(np.int64))          # same order: vslam.cpp:150 takes the first hit
    # RansacFilter
    accepted, ilen = rd.take(np.int32, 2)
    F = rd.take(np.float32, 9)
    mask = rd.take(np.uint8, int(ilen))
    o = oracle.find_fundamental(p1, p2, tent, 8, iters, thr, 1000 + seed)
    assert accepted == 1 and o["best"] >= 0 and ilen == len(tent)
    assert np.array_equal(F.view(np.uint32), o["F"].reshape(-1).view(np.uint32))
    assert np.array_equal(mask, o["mask"])
    cnt = int(rd.take(np.int32, 1)[0])
    score = rd.take(np.float32, 1)[0]
    assert cnt == o["n_inliers"] and score.view(np.uint32) == o["score"].view(np.uint32)
    sets = rd.take(np.int32, iters * 8).reshape(iters, 8)
    assert np.array_equal(sets, oracle.initialize_sets(len(tent), 8, iters, 1000 + seed + 1))   # second seeded call
    F8 = rd.take(np.float32, 9)
    assert np.array_equal(F8.view(np.uint32),
                          oracle.compute_fundamental(p1[tent[:8, 0]], p2[tent[:8, 1]]).reshape(-1).view(np.uint32))


@pytest.mark.gpu
def test_adapter_no_model_leaves_outputs_untouched(tmp_path, oracle):
    """Fewer than min_items matches (UB in the reference): the adapter must leave F empty and inliers as passed."""
    fp = synth.frame_pair(200, 1)
    tent = oracle.match_hamming(fp["d1"], fp["d2"])[:5]
    q = fp["p2"][:4].copy()
    rd = Reader(_run(tmp_path, fp["p1"], fp["p2"], tent, q, 16, 8, 5, 10.0, 3.0))
    k = 200
    rd.take(np.int32, 2); rd.take(np.float32, 2 * k); rd.take(np.float32, 2 * 4)
    for _ in range(4):
        c = int(rd.take(np.int32, 1)[0]); rd.take(np.float32, 2 * c)
    rd.take(np.int32, 2); rd.take(np.int64, k)
    for _ in range(4):
        c = int(rd.take(np.int32, 1)[0]); rd.take(np.int64, c)
    accepted, ilen = rd.take(np.int32, 2)
    assert accepted == 0 and ilen == 0


@pytest.mark.gpu
def test_adapter_reference_kdtree_protocol():
    """tests/test_kdtree.cpp's randomized protocol (integer grid points, ties everywhere) against the adapter."""
    r = subprocess.run([build_harness(), "--protocol", "60"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.splitlines() == ["60 successes out of 60 trials"] * 2
