"""extract_Rt's DISCRETE choices against cv2 over the whole rotation range (reference src/helpers.cpp:18-33).

The reference returns `trace(R_1) < 0 ? R_2 : R_1` with R_1 = U W V^T, R_2 = U W^T V^T. Flipping the sign of one singular
pair (u_i, v_i), i in {0, 1} — which every SVD routine is free to do — swaps R_1 and R_2, so the pick is independent of the
SVD's sign convention only when exactly one of the two traces is negative (always the case for frame-to-frame motions:
one candidate is the motion, trace ~ 3, the other its 180-degree twisted twin, trace ~ -1). This sweep draws 2 000 motions
with rotation angles uniform in [0, pi] plus exactly singular E (F = [t]x, K = I: a zero singular value) and stores cv2's
R, t, so tests/test_oracle_geometry.py can assert the same choice wherever it is convention-independent and quantify the rest.

    python tests/golden/gen_golden_rt_sweep.py        (dev container only: needs cv2)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), HERE]
from gen_golden_geometry import cv_extract_rt  # noqa: E402
from vslam_b200 import synth  # noqa: E402

f32 = np.float32


def rand_rot(rng):
    ax = rng.standard_normal(3)
    ax /= np.linalg.norm(ax)
    a = rng.uniform(0, np.pi)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx, a


def main():
    rng = np.random.default_rng(20261020)
    K = np.array([[525, 0, 640], [0, 525, 360], [0, 0, 1]], f32)
    Fs, Rs, ts, ang = [], [], [], []
    for _ in range(2000):
        R, a = rand_rot(rng)
        F = synth.true_fundamental(R, rng.standard_normal(3)).astype(f32)
        _, _, Rc, tc = cv_extract_rt(F, K)
        Fs.append(F); Rs.append(Rc); ts.append(tc); ang.append(a)
    # exactly rank-2 E: F = [t]x with K = I (the third singular value is exactly 0)
    sF, sR, st = [], [], []
    for tt in ([1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 2, 3], [-2, 1, 4], [3, -1, 2], [0.5, 0.25, 2]):
        tt = np.array(tt, float)
        F = np.array([[0, -tt[2], tt[1]], [tt[2], 0, -tt[0]], [-tt[1], tt[0], 0]], f32)
        _, _, Rc, tc = cv_extract_rt(F, np.eye(3, dtype=f32))
        sF.append(F); sR.append(Rc); st.append(tc)
    path = os.path.join(HERE, "extract_rt_sweep_cv2_4_13.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), K=K, F=np.array(Fs), R=np.array(Rs), t=np.array(ts),
                        angle=np.array(ang, f32), sing_F=np.array(sF), sing_R=np.array(sR), sing_t=np.array(st))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
