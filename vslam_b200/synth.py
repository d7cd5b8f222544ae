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


def projection_scene(n_map: int, k: int, seed: int, nbytes: int = 32, contested: float = 0.3, claimed: float = 0.1):
    """A map + current frame for search by projection (reference src/vslam.cpp:129-161).

    k frame keypoints (a tenth of them in tight clusters, so radius-2 searches return several candidates), n_map
    homogeneous map points: most project within ~1.5 px of a keypoint — `contested` of them onto a keypoint some other
    map point also targets — the rest land elsewhere, outside the image or behind the camera. Every map point has 1-4
    observation descriptors: bit-flipped copies of its target keypoint's descriptor (orb_distance well below 64) or,
    for a fifth of them, unrelated ones. `claimed` of the keypoints already carry a map point id.
    Returns dict(X [n][4], c2 [3][4], pts [k][2], desc [k][nbytes], ids [k], obs_off [n+1], obs_desc [*][nbytes]).
    """
    rng = np.random.default_rng(seed)
    f32 = np.float32
    pts = np.stack([rng.uniform(2, W - 2, k), rng.uniform(2, H - 2, k)], 1)
    nclu = k // 10
    if nclu:   # clusters: copies of other keypoints displaced by < 1.5 px
        src = rng.integers(nclu, k, nclu)
        pts[:nclu] = pts[src] + rng.uniform(-1.0, 1.0, (nclu, 2))
    pts = pts.astype(f32)
    desc = random_descriptors(rng, k, nbytes)
    R, t = default_motion(rng)
    K = np.array([[FOCAL, 0, CX], [0, FOCAL, CY], [0, 0, 1.0]])
    c2 = (K @ np.concatenate([R, t[:, None]], 1)).astype(f32)
    target = rng.integers(0, k, n_map)
    ncont = int(contested * n_map)
    if ncont and n_map > ncont:
        target[:ncont] = target[rng.integers(ncont, n_map, ncont)]
    perm = rng.permutation(n_map)
    target = target[perm]
    uv = pts[target].astype(np.float64) + rng.uniform(-1.4, 1.4, (n_map, 2))
    depth = rng.uniform(3.0, 12.0, n_map)
    kind = rng.uniform(0, 1, n_map)
    uv[kind < 0.10] += rng.uniform(20, 60, (int((kind < 0.10).sum()), 2))        # lands on nothing
    uv[(kind >= 0.10) & (kind < 0.15)] += np.array([W, H])                         # outside the image
    depth[(kind >= 0.15) & (kind < 0.18)] *= -1.0                                  # behind the camera
    xc = np.stack([(uv[:, 0] - CX) / FOCAL * depth, (uv[:, 1] - CY) / FOCAL * depth, depth], 1)
    xw = (xc - t) @ R          # R^T (xc - t)
    X = np.concatenate([xw, np.ones((n_map, 1))], 1).astype(f32)
    X[rng.integers(0, n_map, max(1, n_map // 50))] *= f32(1.5)                     # homogeneous scale != 1
    nobs = rng.integers(1, 5, n_map)
    obs_off = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
    obs_desc = np.zeros((int(obs_off[-1]), nbytes), np.uint8)
    unrelated = rng.uniform(0, 1, n_map) < 0.2
    for i in range(n_map):
        for o in range(obs_off[i], obs_off[i + 1]):
            obs_desc[o] = random_descriptors(rng, 1, nbytes)[0] if unrelated[i] else \
                flip_bits(rng, desc[target[i]][None, :], 40)[0]
    ids = np.full(k, -1, np.int32)
    pre = rng.choice(k, int(claimed * k), replace=False)
    ids[pre] = rng.integers(0, max(n_map, 1), len(pre))
    return {"X": X, "c2": c2, "pts": pts, "desc": desc, "ids": ids, "obs_off": obs_off, "obs_desc": obs_desc}
