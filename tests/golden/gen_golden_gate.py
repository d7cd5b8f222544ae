"""Golden vectors for the reprojection gate after triangulation (reference src/vslam.cpp:186-251) with cv2 4.13.0.

The reference's statements are replayed with the OpenCV entry points they compile to:
    points_4d * c.t()                      -> cv2.gemm(points4, c, GEMM_2_T)                       :192-193
    for (i = 0; i < rows; i += 3) data[i] /= h ...  on the FLAT float data (numpy float32 division)    :201-211
        (the loop bound is the ROW count, so only the first ceil(n/3) points are made non-homogeneous — kept as written)
    reproj.colRange(0, 2) - initial_points -> cv2.subtract                                             :231-232
    d.row(i).dot(d.row(i))                 -> no Python binding exists for cv::Mat::dot; restated from OpenCV's
        dotProd_32f scalar tail (two products and their sum in double), narrowed to f32                :240-242
    gates map_point_ids[i] > 0, re > thresholdSq (4.0, :53), push_back, reproj_error += re1 + re2      :237-251

Sizes sit on both sides of cv::gemm's 100-row switch and of multiples of three. tests/test_oracle_geometry.py requires
vbo_reprojection_gate to reproduce reproj errors bit for bit and the inlier lists exactly.

    python tests/golden/gen_golden_gate.py        (dev container only: needs cv2)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), HERE]
from gen_golden_geometry import cv_triangulate  # noqa: E402
from vslam_b200 import synth  # noqa: E402

f32 = np.float32


def cv_gate(points4, c1, c2, ip1, ip2, ids, thr_sq):
    r1 = cv2.gemm(points4, c1, 1, None, 0, flags=cv2.GEMM_2_T)
    r2 = cv2.gemm(points4, c2, 1, None, 0, flags=cv2.GEMM_2_T)
    f1, f2 = r1.reshape(-1), r2.reshape(-1)
    with np.errstate(all="ignore"):
        for i in range(0, r1.shape[0], 3):                       # rows, not rows * 3: as written
            h1 = f1[i + 2]; f1[i] = f1[i] / h1; f1[i + 1] = f1[i + 1] / h1; f1[i + 2] = 1
            h2 = f2[i + 2]; f2[i] = f2[i] / h2; f2[i + 1] = f2[i + 1] / h2; f2[i + 2] = 1
    d1 = cv2.subtract(np.ascontiguousarray(r1[:, :2]), ip1)
    d2 = cv2.subtract(np.ascontiguousarray(r2[:, :2]), ip2)
    re1 = (d1[:, 0].astype(np.float64) ** 2 + d1[:, 1].astype(np.float64) ** 2).astype(f32)
    re2 = (d2[:, 0].astype(np.float64) ** 2 + d2[:, 1].astype(np.float64) ** 2).astype(f32)
    inl, err = [], 0.0
    for i in range(len(d1)):
        if ids[i] > 0:
            continue
        if re1[i] > f32(thr_sq):
            continue
        if re2[i] > f32(thr_sq):
            continue
        inl.append(i)
        err += float(f32(re1[i] + re2[i]))
    return r1, r2, re1, re2, np.array(inl, np.int32), err


def main():
    rng = np.random.default_rng(20261021)
    K = np.array([[525, 0, 640], [0, 525, 360], [0, 0, 1]], f32)
    out = {}
    for tag, n in (("n1", 1), ("n3", 3), ("n4", 4), ("n98", 98), ("n99", 99), ("n100", 100), ("n101", 101), ("n1500", 1500)):
        R, t = synth.default_motion(rng)
        c1 = np.concatenate([K, np.zeros((3, 1), f32)], 1).astype(f32)
        c2 = (K.astype(np.float64) @ np.concatenate([R, t[:, None]], 1)).astype(f32)
        X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 12, n), np.ones(n)], 1)
        x1 = (c1.astype(np.float64) @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
        x2 = (c2.astype(np.float64) @ X.T).T; x2 = x2[:, :2] / x2[:, 2:]
        p1 = (x1 + rng.normal(0, 0.4, x1.shape)).astype(f32)
        p2 = (x2 + rng.normal(0, 0.4, x2.shape)).astype(f32)
        bad = rng.random(n) < 0.2                                   # gross mismatches: large reprojection error
        p2[bad] += rng.normal(0, 6, (int(bad.sum()), 2)).astype(f32)
        P4 = cv_triangulate(p1, p2, c1, c2)
        ids = np.where(rng.random(n) < 0.25, rng.integers(0, 50, n), -1).astype(np.int32)   # some claimed, some id 0, most -1
        r1, r2, re1, re2, inl, err = cv_gate(P4, c1, c2, p1, p2, ids, 4.0)
        for nm, v in dict(c1=c1, c2=c2, p1=p1, p2=p2, P4=P4, ids=ids, r1=r1, r2=r2, re1=re1, re2=re2, inl=inl,
                          err=np.float64(err)).items():
            out[f"{tag}_{nm}"] = v
    path = os.path.join(HERE, "gate_cv2_4_13.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
