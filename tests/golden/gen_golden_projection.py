"""Golden vectors for the projection step of search-by-projection (reference src/vslam.cpp:126-131) with cv2 4.13.0.

`cv::Mat c2 = K * frame.R_t.rowRange(0, 3)` and `pm.points.rowRange(0, pm.size) * c2.t()` are MatExpr products, i.e.
cv::gemm(K, Rt) and cv::gemm(points, c2, GEMM_2_T). This script issues those two calls for several map sizes on both sides of
OpenCV's small-matrix threshold (100 rows) and stores inputs and outputs; tests/test_oracle_golden.py requires
vbo_project_points to reproduce them bit for bit.

Run from the repo root in the dev container (needs cv2; the GPU box never runs this):
    python tests/golden/gen_golden_projection.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def main():
    rng = np.random.default_rng(20261018)
    K = np.array([[525, 0, 640], [0, 525, 360], [0, 0, 1]], f32)            # src/vslam.cpp:30-32 defaults
    out = {"K": K}
    for tag, n in (("n1", 1), ("n7", 7), ("n99", 99), ("n100", 100), ("n101", 101), ("n640", 640), ("n5000", 5000)):
        rvec = rng.normal(0, 0.03, 3)
        Rt = np.zeros((3, 4), f32)
        Rt[:, :3] = cv2.Rodrigues(rvec)[0].astype(f32)
        Rt[:, 3] = rng.normal(0, 0.3, 3).astype(f32)
        c2 = cv2.gemm(K, Rt, 1, None, 0)                                      # K * R_t.rowRange(0, 3)
        X = np.concatenate([rng.uniform(-6, 6, (n, 2)), rng.uniform(-2, 12, (n, 1)), np.ones((n, 1))], 1).astype(f32)
        X[rng.integers(0, n)] *= f32(1.7)                                     # a row whose w != 1
        P = cv2.gemm(X, c2, 1, None, 0, flags=cv2.GEMM_2_T)                   # points * c2.t()
        out[f"{tag}_Rt"], out[f"{tag}_c2"], out[f"{tag}_X"], out[f"{tag}_P"] = Rt, c2, X, P
    path = os.path.join(HERE, "projection_cv2_4_13.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
