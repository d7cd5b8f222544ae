"""Generate golden vectors for the correspondence path with Python cv2 (OpenCV 4.13.0).

The reference's matcher and RansacFilter are thin sequences of OpenCV calls (src/Frame.cpp:83-85,
src/RansacFilter.cpp:69-140). C++ OpenCV is not installed here, but the cv2 module wraps the same
library, so this script issues the *same OpenCV entry points in the same order* as the reference
(cv::gemm with and without GEMM_1_T, Mat::mul -> cv::multiply, operator/ -> cv::divide,
operator+ -> cv::add, cv::reduce, cv::sum, cv::SVDecomp, BFMatcher::knnMatch) and stores inputs and
outputs. tests/test_oracle_golden.py then requires the C oracle to reproduce them:
  bit-exact : residual e, inlier mask, inlier count, knnMatch indices/distances
  tolerance : (float)cv::sum (summation order inside OpenCV is SIMD-dependent) and the 8-point F
              (cv::SVDecomp result depends on the LAPACK/Jacobi build) — "parity unpinned" steps.

Run from the repo root in the dev container (needs cv2; the GPU box never runs this):
    python tests/golden/gen_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from vslam_b200 import synth  # noqa: E402

f32 = np.float32


def cv_residual(p1, p2, matches, F, thr):
    """src/RansacFilter.cpp:105-140, call for call."""
    n = len(matches)
    x1 = np.ones((3, n), f32)
    x2 = np.ones((3, n), f32)
    x1[0], x1[1] = p1[matches[:, 0], 0], p1[matches[:, 0], 1]
    x2[0], x2[1] = p2[matches[:, 1], 0], p2[matches[:, 1], 1]
    F_x1 = cv2.gemm(F, x1, 1, None, 0)                              # :119  F * x1
    F_t_x2 = cv2.gemm(F, x2, 1, None, 0, flags=cv2.GEMM_1_T)        # :120  F.t() * x2
    s = cv2.reduce(cv2.multiply(x2, F_x1), 0, cv2.REDUCE_SUM)       # :122-123
    with np.errstate(all="ignore"):
        e = cv2.divide(cv2.multiply(s, s), cv2.multiply(F_x1[0:1], F_x1[0:1]))   # :126, as parsed
        e = cv2.add(e, cv2.multiply(F_x1[1:2], F_x1[1:2]))
        e = cv2.add(e, cv2.multiply(F_t_x2[0:1], F_t_x2[0:1]))
        e = cv2.add(e, cv2.multiply(F_t_x2[1:2], F_t_x2[1:2]))
    e = e.reshape(-1)
    mask = (e <= f32(thr)).astype(np.uint8)                          # :130
    score = f32(cv2.sumElems(e.reshape(1, -1))[0])                   # :138
    return e, mask, int(mask.sum()), score


def cv_compute_fundamental(p1s, p2s):
    """src/RansacFilter.cpp:69-103 with cv2.SVDecomp."""
    A = np.empty((8, 9), f32)
    u1, v1, u2, v2 = p1s[:, 0], p1s[:, 1], p2s[:, 0], p2s[:, 1]
    A[:, 0], A[:, 1], A[:, 2] = u2 * u1, u2 * v1, u2
    A[:, 3], A[:, 4], A[:, 5] = v2 * u1, v2 * v1, v2
    A[:, 6], A[:, 7], A[:, 8] = u1, v1, 1
    w, u, vt = cv2.SVDecomp(A.copy(), flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
    F = vt[8].reshape(3, 3).copy()
    w, u, vt = cv2.SVDecomp(F.copy(), flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
    w = w.reshape(-1).copy()
    w[2] = 0
    return A, F, cv2.gemm(cv2.gemm(u, np.diag(w).astype(f32), 1, None, 0), vt, 1, None, 0)


def main():
    out = {}
    rng = np.random.default_rng(20261018)

    # ---- residual: realistic F's (8-point solutions of random minimal samples) --------------
    fp = synth.frame_pair(600, seed=3)
    gt = fp["gt"]
    keep = np.nonzero(gt >= 0)[0]
    # tentative matches: mostly correct, some wrong
    matches = np.stack([keep, gt[keep]], 1).astype(np.int32)
    wrong = rng.choice(len(matches), 40, replace=False)
    matches[wrong, 1] = rng.integers(0, 600, 40)
    Fs, es, masks, cnts, scores = [], [], [], [], []
    for h in range(24):
        sel = rng.choice(len(matches), 8, replace=False)
        _, _, F = cv_compute_fundamental(fp["p1"][matches[sel, 0]], fp["p2"][matches[sel, 1]])
        F = np.ascontiguousarray(F, f32)
        e, mask, cnt, score = cv_residual(fp["p1"], fp["p2"], matches, F, 10.0)
        Fs.append(F); es.append(e); masks.append(mask); cnts.append(cnt); scores.append(score)
    # adversarial F's: zero rows (a0 == 0 -> x/0, 0/0), huge and tiny scales, negative zero
    adv = [np.zeros((3, 3), f32), np.array([[0, 0, 0], [1e-3, 2e-3, -1], [3e-6, 1e-6, 0.5]], f32),
           np.eye(3, dtype=f32) * f32(1e-20), np.eye(3, dtype=f32) * f32(1e18),
           np.array([[-0.0, 0, 0], [0, -0.0, 1], [0, -1, 0]], f32), rng.standard_normal((3, 3)).astype(f32)]
    for F in adv:
        e, mask, cnt, score = cv_residual(fp["p1"], fp["p2"], matches, F, 10.0)
        Fs.append(F); es.append(e); masks.append(mask); cnts.append(cnt); scores.append(score)
    out.update(res_p1=fp["p1"], res_p2=fp["p2"], res_matches=matches, res_F=np.stack(Fs), res_e=np.stack(es),
               res_mask=np.stack(masks), res_cnt=np.array(cnts, np.int32), res_score=np.array(scores, f32),
               res_thr=f32(10.0))

    # ---- 8-point solve: A and F from cv2.SVDecomp (tolerance only) -----------------------------
    As, F0s, F2s, P1s, P2s = [], [], [], [], []
    for h in range(32):
        sel = rng.choice(keep, 8, replace=False)
        p1s, p2s = fp["p1"][sel], fp["p2"][gt[sel]]
        A, F0, F2 = cv_compute_fundamental(p1s, p2s)
        As.append(A); F0s.append(F0); F2s.append(F2); P1s.append(p1s); P2s.append(p2s)
    out.update(fm_A=np.stack(As), fm_F0=np.stack(F0s), fm_F2=np.stack(F2s), fm_p1=np.stack(P1s),
               fm_p2=np.stack(P2s))

    # ---- small gemm rules (3x3 * 3x3 fp32 path used by U*diag(D)*Vt) ---------------------------
    Ma = rng.standard_normal((16, 3, 3)).astype(f32)
    Mb = rng.standard_normal((16, 3, 3)).astype(f32)
    out.update(g3_a=Ma, g3_b=Mb, g3_c=np.stack([cv2.gemm(a, b, 1, None, 0) for a, b in zip(Ma, Mb)]))

    # ---- knnMatch: tie order and distances ------------------------------------------------------
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    d2 = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    d1 = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    d2[400:460] = d2[0:60]            # exact duplicate train rows -> distance ties
    d1[:60] = synth.flip_bits(rng, d2[:60], 6)
    d1[60:80] = d2[100:120]           # zero-distance matches
    coarse2 = (rng.integers(0, 2, (64, 32), dtype=np.uint8) * 255)   # few distinct distances -> many ties
    coarse1 = (rng.integers(0, 2, (48, 32), dtype=np.uint8) * 255)
    for name, (q, t) in dict(knn=(d1, d2), knnc=(coarse1, coarse2)).items():
        mm = bf.knnMatch(q, t, k=2)
        idx = np.array([[m[0].trainIdx, m[1].trainIdx] for m in mm], np.int32)
        dist = np.array([[m[0].distance, m[1].distance] for m in mm], f32)
        keepm = np.array([m[0].distance < m[1].distance * 0.7 for m in mm], np.uint8)   # src/Frame.cpp:91
        out.update({f"{name}_d1": q, f"{name}_d2": t, f"{name}_idx": idx, f"{name}_dist": dist, f"{name}_keep": keepm})

    path = os.path.join(HERE, "correspondence_cv2_4_13.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
