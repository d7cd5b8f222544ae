"""extract_Rt / triangulate restatements (reference src/helpers.cpp) against cv2 goldens.
E = K^T F K: bit-exact. R, t, triangulated points: tolerance (the SVDs are build-defined, 'parity unpinned')."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geometry_cv2_4_13.npz"))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_essential_bit_exact_vs_cv2(oracle):
    for F, E in zip(G["F"], G["E"]):
        assert np.array_equal(_bits(oracle.essential(F, G["K"])), _bits(E))


def test_extract_rt_vs_cv2_tolerance(oracle):
    # tolerance written here: 2e-4 absolute on the entries of R (|R_ij| <= 1) and of the unit vector t
    for F, R, t, Rt in zip(G["F"], G["R"], G["t"], G["R_true"]):
        Ro, to = oracle.extract_rt(F, G["K"])
        assert np.abs(Ro.astype(np.float64) @ Ro.T.astype(np.float64) - np.eye(3)).max() < 1e-5
        assert np.linalg.det(Ro.astype(np.float64)) > 0.999
        assert abs(float(np.linalg.norm(to.astype(np.float64))) - 1.0) < 1e-6 and to[2] >= 0
        assert np.abs(Ro - R).max() < 2e-4 and np.abs(to - t).max() < 2e-4


def test_extract_rt_recovers_known_motion(oracle):
    for F, Rt, tt in zip(G["F"][0::3], G["R_true"][0::3], G["t_true"][0::3]):   # the unperturbed F's
        Ro, to = oracle.extract_rt(F, G["K"])
        assert np.abs(Ro - Rt).max() < 2e-3
        tn = tt / np.linalg.norm(tt)
        assert min(np.abs(to - tn).max(), np.abs(to + tn).max()) < 5e-3


def test_triangulate_vs_cv2_and_ground_truth(oracle):
    P = oracle.triangulate(G["tri_p1"], G["tri_p2"], G["tri_c1"], G["tri_c2"])
    assert np.array_equal(P[:, 3], np.ones(len(P), np.float32))
    ref, X = G["tri_P4"], G["tri_X"]
    # tolerance written here: 2e-3 relative to the point's depth against cv2's fp32 SVD; and both lie equally close to
    # the true points (0.3 px noise dominates)
    assert (np.abs(P[:, :3] - ref[:, :3]).max(1) / np.abs(X[:, 2])).max() < 2e-3
    eo = np.linalg.norm(P[:, :3] - X[:, :3], axis=1)
    er = np.linalg.norm(ref[:, :3] - X[:, :3], axis=1)
    assert np.median(eo) < 0.5 and abs(np.median(eo) - np.median(er)) < 1e-2
