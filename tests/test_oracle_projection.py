"""CPU checks of the search-by-projection restatement (oracle) — reference src/vslam.cpp:129-161.

The C function is compared with the same loop written out in Python over the oracle's primitives (projection pinned
against cv2 in test_oracle_golden.py, kd-tree radius search pinned against the reference's own KDTree.cpp in
test_oracle_vs_ref.py, numpy popcount for orb_distance), and with the reference's own radius_search where oracle/_ref
was built."""
import numpy as np
import pytest

from vslam_b200 import synth


def _python_loop(oracle, s, W, H, radius, thr, radius_fn=None):
    X, c2, pts, desc = s["X"], s["c2"], s["pts"], s["desc"]
    pr = oracle.project_points(X, c2)
    pre = oracle.kdtree_build(pts)
    ids = s["ids"].copy()
    assign = np.full(len(X), -1, np.int32)
    f = np.float32
    for i in range(len(X)):
        with np.errstate(all="ignore"):
            x, y = f(pr[i, 0] / pr[i, 2]), f(pr[i, 1] / pr[i, 2])
        if not (x >= 0 and x < W and y >= 0 and y < H):
            continue
        hits = radius_fn(np.array([x, y], f)) if radius_fn else oracle.kdtree_radius(pts, pre, (x, y), radius)[0]
        obs = s["obs_desc"][s["obs_off"][i]:s["obs_off"][i + 1]]
        for idx in hits:
            if ids[idx] >= 0:
                continue
            d = min((int(np.unpackbits(desc[idx] ^ o).sum()) for o in obs), default=0xffffffff)
            if d < thr:
                ids[idx] = i
                assign[i] = idx
                break
    return assign, ids


@pytest.mark.parametrize("n_map,k,seed", [(400, 600, 1), (60, 300, 2), (1500, 200, 3)])
def test_oracle_search_by_projection_equals_python_loop(oracle, n_map, k, seed):
    s = synth.projection_scene(n_map, k, seed)
    a, ids, xy, inv = oracle.search_by_projection(s["X"], s["c2"], 1280, 720, s["pts"], s["desc"], s["ids"], s["obs_off"],
                                                  s["obs_desc"], 2.0, 64)
    pa, pids = _python_loop(oracle, s, 1280, 720, 2.0, 64)
    assert np.array_equal(a, pa) and np.array_equal(ids, pids)
    claimed = a[a >= 0]
    assert len(claimed) == len(set(claimed.tolist())) and len(claimed) > 0
    assert (s["ids"][claimed] < 0).all()   # only keypoints that were free at entry get claimed


def test_oracle_search_by_projection_with_reference_radius_search(oracle, ref):
    """Candidate order comes from the reference's own radius_search(frame_kdtree) (src/KDTree.cpp:145-171)."""
    s = synth.projection_scene(500, 700, 5)
    a, ids, _, _ = oracle.search_by_projection(s["X"], s["c2"], 1280, 720, s["pts"], s["desc"], s["ids"], s["obs_off"],
                                               s["obs_desc"], 2.0, 64)

    def ref_radius(q):
        off, hits = ref.frame_kdtree_radius(s["pts"], np.ascontiguousarray(q[None, :]), 2.0, cap=256)
        return hits[off[0]:off[1]]

    pa, pids = _python_loop(oracle, s, 1280, 720, 2.0, 64, radius_fn=ref_radius)
    assert np.array_equal(a, pa) and np.array_equal(ids, pids)
