"""Deterministic synthetic keypoint / descriptor sets for the correspondence path.

Shapes follow SURVEY.md §8d (configs C1-C5): a 1280x720 image, the reference's default intrinsics
K = [[525,0,640],[0,525,360],[0,0,1]] (src/vslam.cpp:30,32 of the reference), real-valued pixel
coordinates (so KD-tree split coordinates are distinct), 256-bit binary descriptors (ORB-sized,
src/Frame.cpp:68) or 128-d unit-norm float descriptors (BASELINE config 3).

Everything is numpy + an explicit seed; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

W, H = 1280.0, 720.0
FOCAL = 525.0
CX, CY = 640.0, 360.0


def _rot(rx: float, ry: float, rz: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def default_motion(rng: np.random.Generator | None = None):
    """Small rotation + forward/vertical translation (the reference's residual is only
    meaningful for such motions, not for sideways translation)."""
    if rng is None:
        return _rot(0.004, -0.006, 0.003), np.array([0.02, 0.10, 0.35])
    r = rng.normal(0, 0.004, 3)
    t = np.array([rng.normal(0, 0.02), rng.normal(0.08, 0.03), rng.uniform(0.2, 0.5)])
    return _rot(*r), t


def true_fundamental(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """F with x2^T F x1 = 0 for X2 = R X1 + t."""
    K = np.array([[FOCAL, 0, CX], [0, FOCAL, CY], [0, 0, 1.0]])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Kinv = np.linalg.inv(K)
    F = Kinv.T @ tx @ R @ Kinv
    return F / np.linalg.norm(F)


def random_descriptors(rng: np.random.Generator, n: int, nbytes: int = 32) -> np.ndarray:
    return rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)


def flip_bits(rng: np.random.Generator, desc: np.ndarray, max_flips: int = 20) -> np.ndarray:
    """Copy of desc with 0..max_flips random bit positions toggled per row."""
    n, nbytes = desc.shape
    out = desc.copy()
    nflip = rng.integers(0, max_flips + 1, size=n)
    pos = rng.integers(0, nbytes * 8, size=(n, max_flips))
    for j in range(max_flips):
        act = nflip > j
        byte = pos[:, j] >> 3
        bit = (1 << (pos[:, j] & 7)).astype(np.uint8)
        rows = np.nonzero(act)[0]
        out[rows, byte[rows]] ^= bit[rows]
    return out


def _advance(rng, pts, depth, R, t, noise_px, outlier_frac):
    """Project frame-1 keypoints (pixel + depth) into the next frame; returns pts2, depth2, inlier flag."""
    n = pts.shape[0]
    X = np.stack([(pts[:, 0] - CX) / FOCAL * depth, (pts[:, 1] - CY) / FOCAL * depth, depth], 1)
    X2 = X @ R.T + t
    z = X2[:, 2]
    u = FOCAL * X2[:, 0] / z + CX + rng.normal(0, noise_px, n)
    v = FOCAL * X2[:, 1] / z + CY + rng.normal(0, noise_px, n)
    ok = (z > 0.5) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
    ok &= rng.random(n) >= outlier_frac
    pts2 = np.stack([u, v], 1)
    nb = int((~ok).sum())
    pts2[~ok] = np.stack([rng.uniform(0, W, nb), rng.uniform(0, H, nb)], 1)
    depth2 = np.where(ok, z, rng.uniform(4.0, 12.0, n))
    return pts2, depth2, ok


def frame_pair(k: int, seed: int, noise_px: float = 0.5, outlier_frac: float = 0.3, nbytes: int = 32,
               shuffle: bool = True):
    """Config C1/C2: two frames of k keypoints with a known ground-truth F.

    Returns dict(p1, d1, p2, d2 (float32 / uint8, C-contiguous), F_true, gt (index in frame 2 of each
    frame-1 keypoint, -1 for outliers)).
    """
    rng = np.random.default_rng(seed)
    p1 = np.stack([rng.uniform(0, W, k), rng.uniform(0, H, k)], 1)
    depth = rng.uniform(4.0, 12.0, k)
    R, t = default_motion()
    p2, _, ok = _advance(rng, p1, depth, R, t, noise_px, outlier_frac)
    d1 = random_descriptors(rng, k, nbytes)
    d2 = flip_bits(rng, d1)
    nb = int((~ok).sum())
    d2[~ok] = random_descriptors(rng, nb, nbytes)
    perm = rng.permutation(k) if shuffle else np.arange(k)
    inv = np.empty(k, np.int64)
    inv[perm] = np.arange(k)
    gt = np.where(ok, inv, -1)
    return dict(p1=np.ascontiguousarray(p1, np.float32), d1=np.ascontiguousarray(d1),
                p2=np.ascontiguousarray(p2[perm], np.float32), d2=np.ascontiguousarray(d2[perm]),
                F_true=true_fundamental(R, t), gt=gt)


def frame_pair_float(k: int, seed: int, dim: int = 128, outlier_frac: float = 0.3, sigma: float = 0.05):
    """Config C3: unit-norm fp32 Gaussian descriptors; inlier copy + N(0, sigma) re-normalised."""
    base = frame_pair(k, seed, outlier_frac=outlier_frac, nbytes=1, shuffle=True)
    rng = np.random.default_rng(seed + 7919)
    d1 = rng.standard_normal((k, dim)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 = np.empty_like(d1)
    gt = base["gt"]
    fresh = rng.standard_normal((k, dim)).astype(np.float32)
    fresh /= np.linalg.norm(fresh, axis=1, keepdims=True)
    d2[:] = fresh
    src = np.nonzero(gt >= 0)[0]
    pert = d1[src] + rng.normal(0, sigma, (src.size, dim)).astype(np.float32)
    pert /= np.linalg.norm(pert, axis=1, keepdims=True)
    d2[gt[src]] = pert
    base["d1"] = np.ascontiguousarray(d1, np.float32)
    base["d2"] = np.ascontiguousarray(d2, np.float32)
    return base


def sequence(nframes: int, k: int, seed: int, noise_px: float = 0.5, outlier_frac: float = 0.3,
             nbytes: int = 32):
    """Config C4: a smooth forward trajectory. Returns pts [nframes,k,2] float32, desc [nframes,k,nbytes] uint8.

    Frame i+1 is frame i advanced by a small random motion: surviving keypoints keep their descriptor
    with 0-20 flipped bits, the rest are replaced by fresh keypoints; order is shuffled per frame.
    """
    rng = np.random.default_rng(seed)
    pts = np.empty((nframes, k, 2), np.float32)
    desc = np.empty((nframes, k, nbytes), np.uint8)
    p = np.stack([rng.uniform(0, W, k), rng.uniform(0, H, k)], 1)
    depth = rng.uniform(4.0, 12.0, k)
    d = random_descriptors(rng, k, nbytes)
    pts[0], desc[0] = p, d
    for i in range(1, nframes):
        R, t = default_motion(rng)
        p2, depth2, ok = _advance(rng, p, depth, R, t, noise_px, outlier_frac)
        d2 = flip_bits(rng, d)
        nb = int((~ok).sum())
        d2[~ok] = random_descriptors(rng, nb, nbytes)
        perm = rng.permutation(k)
        p, depth, d = p2[perm], depth2[perm], d2[perm]
        pts[i], desc[i] = p, d
    return pts, desc


def correspondences(m: int, seed: int, outlier_frac: float = 0.3):
    """Config C5: m correspondences as float32 [m,4] rows (x1,y1,x2,y2) from the C1 generator."""
    fp = frame_pair(m, seed, outlier_frac=outlier_frac, nbytes=1, shuffle=False)
    return np.ascontiguousarray(np.concatenate([fp["p1"], fp["p2"]], 1), np.float32)
