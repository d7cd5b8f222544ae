"""Golden vectors for extract_Rt / triangulate (reference src/helpers.cpp) with cv2 4.13.0: the same OpenCV entry points in
the same order (cv::gemm with/without GEMM_1_T, cv::SVD::compute -> cv2.SVDecomp, cv::determinant, cv::norm).
E = K^T F K is required bit for bit of the oracle; R, t and the triangulated points (SVD results) to a tolerance.

    python tests/golden/gen_golden_geometry.py        (dev container only: needs cv2)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from vslam_b200 import synth  # noqa: E402

f32 = np.float32


def cv_extract_rt(F, K):
    """src/helpers.cpp:3-35, call for call."""
    T = cv2.gemm(K, F, 1, None, 0, flags=cv2.GEMM_1_T)      # K.t() * fundamental
    E = cv2.gemm(T, K, 1, None, 0)                          # ... * K
    D, U, Vt = cv2.SVDecomp(E)                              # cv::SVD::compute(E, D, U, V_t)
    t = U[:, 2:3].copy()
    t = (t / cv2.norm(t)).astype(f32)
    W = np.zeros((3, 3), f32); W[0, 1] = -1; W[1, 0] = 1; W[2, 2] = 1
    R1 = cv2.gemm(cv2.gemm(U, W, 1, None, 0), Vt, 1, None, 0)
    if cv2.determinant(R1) < 0:
        R1 = -R1
    R2 = cv2.gemm(cv2.gemm(U, W.T.copy(), 1, None, 0), Vt, 1, None, 0)
    if cv2.determinant(R2) < 0:
        R2 = -R2
    R = R2 if (R1[0, 0] + R1[1, 1] + R1[2, 2] < 0) else R1
    if t[2, 0] < 0:
        t = -t
    return T, E, R, t[:, 0]


def cv_triangulate(p1, p2, c1, c2):
    """src/helpers.cpp:37-80."""
    out = np.zeros((len(p1), 4), f32)
    for i in range(len(p1)):
        A = np.stack([p1[i, 0] * c1[2] - c1[0], p1[i, 1] * c1[2] - c1[1],
                      p2[i, 0] * c2[2] - c2[0], p2[i, 1] * c2[2] - c2[1]]).astype(f32)
        _, _, Vt = cv2.SVDecomp(A, flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
        out[i] = [Vt[3, 0] / Vt[3, 3], Vt[3, 1] / Vt[3, 3], Vt[3, 2] / Vt[3, 3], 1]
    return out


def main():
    rng = np.random.default_rng(20261019)
    K = np.array([[525, 0, 640], [0, 525, 360], [0, 0, 1]], f32)
    Fs, Ts, Es, Rs, ts, Rtrue, ttrue = [], [], [], [], [], [], []
    for i in range(24):
        R, t = synth.default_motion(rng)
        F = synth.true_fundamental(R, t)
        F = (F / np.linalg.norm(F)).astype(f32)
        if i % 3 == 1:
            F = (F + rng.normal(0, 1e-9, (3, 3))).astype(f32)      # a slightly perturbed (estimated) F
        T, E, Rc, tc = cv_extract_rt(F, K)
        Fs.append(F); Ts.append(T); Es.append(E); Rs.append(Rc); ts.append(tc); Rtrue.append(R); ttrue.append(t)
    # triangulation: two cameras, points in front of both, 0.3 px noise
    R, t = synth.default_motion(rng)
    c1 = np.concatenate([K, np.zeros((3, 1), f32)], 1).astype(f32)
    c2 = (K.astype(np.float64) @ np.concatenate([R, t[:, None]], 1)).astype(f32)
    n = 400
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 12, n), np.ones(n)], 1)
    x1 = (c1.astype(np.float64) @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    x2 = (c2.astype(np.float64) @ X.T).T; x2 = x2[:, :2] / x2[:, 2:]
    p1 = (x1 + rng.normal(0, 0.3, x1.shape)).astype(f32)
    p2 = (x2 + rng.normal(0, 0.3, x2.shape)).astype(f32)
    P4 = cv_triangulate(p1, p2, c1, c2)
    path = os.path.join(HERE, "geometry_cv2_4_13.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), K=K, F=np.array(Fs), T=np.array(Ts), E=np.array(Es),
                        R=np.array(Rs), t=np.array(ts), R_true=np.array(Rtrue), t_true=np.array(ttrue),
                        tri_c1=c1, tri_c2=c2, tri_p1=p1, tri_p2=p2, tri_X=X.astype(f32), tri_P4=P4)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
