"""End-to-end goldens for RansacFilter::find_fundamental run ENTIRELY through Python cv2 (OpenCV 4.13.0).

Why: the 8-point solve (cv::SVDecomp x2, reference src/RansacFilter.cpp:94,98) is the one step of the path whose
arithmetic this build cannot pin bit-for-bit (cv2's SVD is LAPACK sgesdd here; the oracle defines an fp64 QR + Jacobi
sequence instead). This script measures what that costs where it matters — the OUTPUT of
find_fundamental — by replaying the reference loop call for call with cv2:

    for every hypothesis h (sample sets = std::mt19937 + libstdc++ uniform_int_distribution, which ARE pinned:
    tests/test_oracle_vs_ref.py::test_sample_sets_match_libstdcxx):
        compute_fundamental      :69-103   cv2.SVDecomp, cv2.gemm
        compute_fundamental_residual :105-140  cv2.gemm / multiply / divide / add / reduce / sumElems
        strict sequential update :59       n > best_n || (n == best_n && score > best_score)

on >= 50 seeded problems of BASELINE configs 1 (2 000 kpts, H = 100, thr = 10 — src/vslam.cpp:19) and 2 (5 000 kpts,
H = 1 024). Stored per problem: cv2's winner index, inlier count, score, F, inlier mask, and every hypothesis' inlier
count. tests/test_oracle_golden_e2e.py compares the C oracle (hence, bit for bit, the GPU path) with these.

Inputs are regenerated from vslam_b200.synth by seed; a SHA-1 of the exact input bytes is stored so drift is caught.
Tentative matches come from cv2.BFMatcher.knnMatch + the ratio test exactly as src/Frame.cpp:83-95.

Run from the repo root in the dev container (needs cv2; the GPU box never runs this):
    python tests/golden/gen_golden_e2e.py
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from gen_golden import cv_compute_fundamental, cv_residual  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from vslam_b200 import synth  # noqa: E402

f32 = np.float32

PROBLEMS = [(2000, 100, s) for s in range(30)] + [(5000, 1024, s) for s in range(24)]
THR = 10.0
RATIO = 0.7


def ransac_seed(k, s):
    return 7000 + 13 * s + k


def cv_tentative(d1, d2):
    """src/Frame.cpp:83-95."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    mm = bf.knnMatch(d1, d2, k=2)
    return np.array([(m[0].queryIdx, m[0].trainIdx) for m in mm if m[0].distance < m[1].distance * RATIO], np.int32)


def input_digest(fp, tent):
    h = hashlib.sha1()
    for a in (fp["p1"], fp["p2"], fp["d1"], fp["d2"], tent):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def cv_find_fundamental(p1, p2, tent, sets):
    """src/RansacFilter.cpp:36-67 with every OpenCV call made through cv2."""
    best_n, best_score, best = 0, f32(0), -1
    best_F, best_mask = np.zeros((3, 3), f32), np.zeros(len(tent), np.uint8)
    cnts = np.zeros(len(sets), np.int32)
    scores = np.zeros(len(sets), f32)
    Fs = np.zeros((len(sets), 9), f32)
    for h, st in enumerate(sets):
        mm = tent[st]
        _, _, F = cv_compute_fundamental(p1[mm[:, 0]], p2[mm[:, 1]])
        F = np.ascontiguousarray(F, f32)
        _, mask, n, score = cv_residual(p1, p2, tent, F, THR)
        cnts[h], scores[h], Fs[h] = n, score, F.reshape(-1)
        if n > best_n or (n == best_n and score > best_score):      # :59
            best_n, best_score, best, best_F, best_mask = n, score, h, F, mask
    return dict(best=best, n=best_n, score=best_score, F=best_F, mask=best_mask, cnts=cnts, scores=scores, Fs=Fs)


def main():
    orc = Oracle()
    out = dict(problems=np.array(PROBLEMS, np.int32), thr=f32(THR), ratio=np.float64(RATIO))
    digests = []
    for i, (k, H, s) in enumerate(PROBLEMS):
        fp = synth.frame_pair(k, s)
        tent = cv_tentative(fp["d1"], fp["d2"])
        assert np.array_equal(tent, orc.match_hamming(fp["d1"], fp["d2"], RATIO))      # the matcher IS pinned
        sets = orc.initialize_sets(len(tent), 8, H, ransac_seed(k, s))
        r = cv_find_fundamental(fp["p1"], fp["p2"], tent, sets)
        digests.append(input_digest(fp, tent))
        out[f"best_{i}"] = np.int32(r["best"])
        out[f"n_{i}"] = np.int32(r["n"])
        out[f"score_{i}"] = f32(r["score"])
        out[f"F_{i}"] = r["F"]
        out[f"mask_{i}"] = np.packbits(r["mask"])
        out[f"m_{i}"] = np.int32(len(tent))
        out[f"cnts_{i}"] = r["cnts"]
        print(f"problem {i}: k={k} H={H} seed={s} m={len(tent)} cv2 best={r['best']} n={r['n']}", flush=True)
    out["digests"] = np.array(digests)
    path = os.path.join(HERE, "find_fundamental_cv2_4_13.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
