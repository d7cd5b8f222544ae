"""ctypes bindings for the CHECKER libraries (oracle/_build/liboracle.so, oracle/_ref/libvbref.so).

Test infrastructure. Imported only from tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build_oracle(out: str = "_build", march: str | None = None) -> str:
    args = ["make", "-C", ORACLE_DIR, f"OUT={out}"]
    if march:
        args.append(f"MARCH={march}")
    subprocess.run(args, check=True, capture_output=True)
    return os.path.join(ORACLE_DIR, out, "liboracle.so")


def _opt(arr, ptr_t):
    return None if arr is None else arr.ctypes.data_as(ptr_t)


class Oracle:
    """The C restatement (oracle/vb_oracle.c)."""

    def __init__(self, path: str | None = None):
        if path is None:
            path = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
            if not os.path.exists(path):
                build_oracle()
        L = self.lib = C.CDLL(path)
        L.vbo_score_sum.restype = C.c_double
        L.vbo_score_sum.argtypes = [_f32p, C.c_int]
        L.vbo_initialize_sets.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, _i32p]
        L.vbo_null_vector_8x9.argtypes = [_f32p, _f32p]
        L.vbo_svd3x3.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.vbo_compute_fundamental.argtypes = [_f32p, _f32p, _f32p]
        L.vbo_residual_one.restype = C.c_float
        L.vbo_residual_one.argtypes = [_f32p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.vbo_compute_fundamental_residual.argtypes = [_f32p, _f32p, _i32p, C.c_int, _f32p, C.c_float, _u8p, _f32p,
                                                       C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.vbo_find_fundamental.restype = C.c_int
        L.vbo_find_fundamental.argtypes = [_f32p, _f32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint32,
                                           _f32p, _u8p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vbo_knn2_hamming.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, _i32p, _i32p]
        L.vbo_ratio_keep.restype = C.c_int
        L.vbo_ratio_keep.argtypes = [C.c_int, C.c_int, C.c_double]
        L.vbo_match_hamming.restype = C.c_int
        L.vbo_match_hamming.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_double, _i32p]
        L.vbo_knn2_l2f.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p]
        L.vbo_match_l2f.restype = C.c_int
        L.vbo_match_l2f.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_double, _i32p]
        L.vbo_match_features.restype = C.c_int
        L.vbo_match_features.argtypes = [_f32p, _u8p, C.c_int, _f32p, _u8p, C.c_int, C.c_int, C.c_double, C.c_int,
                                         C.c_int, C.c_float, C.c_uint32, _i32p, _f32p, C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]
        L.vbo_kdtree_build.argtypes = [_f32p, C.c_int, _i32p]
        L.vbo_kdtree_height.restype = C.c_int
        L.vbo_kdtree_nearest.restype = C.c_int
        L.vbo_kdtree_nearest.argtypes = [_f32p, _i32p, C.c_int, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.vbo_kdtree_knn.restype = C.c_int
        L.vbo_kdtree_knn.argtypes = [_f32p, _i32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, _i32p, _f32p]
        L.vbo_kdtree_radius.restype = C.c_int
        L.vbo_kdtree_radius.argtypes = [_f32p, _i32p, C.c_int, C.c_float, C.c_float, C.c_float, _i32p, C.c_int]
        L.vbo_orb_distance.restype = C.c_uint32
        L.vbo_orb_distance.argtypes = [_u8p, _u8p, C.c_int, C.c_int]
        L.vbo_project_points.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.vbo_search_by_projection.restype = C.c_int
        L.vbo_search_by_projection.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _f32p, _i32p, C.c_int, _u8p, C.c_int,
                                               _i32p, _i32p, _u8p, C.c_float, C.c_uint32, _i32p, _f32p, _u8p]
        L.vbo_essential.argtypes = [_f32p, _f32p, _f32p]
        L.vbo_extract_rt.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.vbo_null_vector_4x4.argtypes = [_f32p, _f32p]
        L.vbo_triangulate.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p, _f32p]
        L.vbo_reprojection_gate.restype = C.c_int
        L.vbo_reprojection_gate.argtypes = [_f32p, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_float, _f32p, _f32p, _i32p,
                                            C.POINTER(C.c_double)]
        L.vbo_compute_fundamental_hartley.argtypes = [_f32p, _f32p, _f32p]
        L.vbo_sampson_one.restype = C.c_float
        L.vbo_sampson_one.argtypes = [_f32p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.vbo_find_fundamental_ex.restype = C.c_int
        L.vbo_find_fundamental_ex.argtypes = [_f32p, _f32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint32, C.c_uint,
                                              _f32p, _u8p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.vbo_pairs_run.restype = C.c_long
        L.vbo_pairs_run.argtypes = [_f32p, _u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_float,
                                    C.c_uint32, C.c_int, C.POINTER(C.c_int)]

    # --- sampling / solve / residual ---
    def score_sum(self, e):
        e = np.ascontiguousarray(e, np.float32)
        return self.lib.vbo_score_sum(e, e.size)

    def initialize_sets(self, n_matches, min_items, iters, seed):
        sets = np.zeros((iters, 8), np.int32)
        self.lib.vbo_initialize_sets(n_matches, min_items, iters, seed, sets)
        return sets

    def null_vector(self, A):
        f = np.zeros(9, np.float32)
        self.lib.vbo_null_vector_8x9(np.ascontiguousarray(A, np.float32), f)
        return f

    def svd3(self, F):
        U, D, Vt = np.zeros((3, 3), np.float32), np.zeros(3, np.float32), np.zeros((3, 3), np.float32)
        self.lib.vbo_svd3x3(np.ascontiguousarray(F, np.float32), U, D, Vt)
        return U, D, Vt

    def compute_fundamental(self, p1set, p2set):
        F = np.zeros((3, 3), np.float32)
        self.lib.vbo_compute_fundamental(np.ascontiguousarray(p1set, np.float32),
                                         np.ascontiguousarray(p2set, np.float32), F)
        return F

    def residual(self, p1, p2, matches, F, thr):
        m = len(matches)
        mask, e = np.zeros(max(m, 1), np.uint8), np.zeros(max(m, 1), np.float32)
        n, s = C.c_int(), C.c_float()
        self.lib.vbo_compute_fundamental_residual(p1, p2, np.ascontiguousarray(matches, np.int32), m,
                                                  np.ascontiguousarray(F, np.float32).reshape(-1), thr, mask, e,
                                                  C.byref(n), C.byref(s))
        return mask[:m], e[:m], n.value, np.float32(s.value)

    def find_fundamental(self, p1, p2, matches, min_items, iters, thr, seed, want_all=False):
        m = len(matches)
        matches = np.ascontiguousarray(matches, np.int32)
        F, mask = np.zeros((3, 3), np.float32), np.zeros(max(m, 1), np.uint8)
        n, s, b = C.c_int(), C.c_float(), C.c_int()
        sets = np.zeros((max(iters, 1), 8), np.int32)
        Fall = np.zeros((max(iters, 1), 9), np.float32)
        cnt = np.zeros(max(iters, 1), np.int32)
        sc = np.zeros(max(iters, 1), np.float32)
        rc = self.lib.vbo_find_fundamental(p1, p2, matches, m, min_items, iters, thr, seed, F, mask, C.byref(n),
                                           C.byref(s), C.byref(b), sets.ctypes.data, Fall.ctypes.data,
                                           cnt.ctypes.data, sc.ctypes.data)
        out = dict(rc=rc, F=F, mask=mask[:m], n_inliers=n.value, score=np.float32(s.value), best=b.value)
        if want_all:
            out.update(sets=sets[:iters], F_all=Fall[:iters], cnt_all=cnt[:iters], score_all=sc[:iters])
        return out

    def compute_fundamental_hartley(self, p1set, p2set):
        F = np.zeros((3, 3), np.float32)
        self.lib.vbo_compute_fundamental_hartley(np.ascontiguousarray(p1set, np.float32), np.ascontiguousarray(p2set, np.float32), F)
        return F

    def find_fundamental_ex(self, p1, p2, matches, min_items, iters, thr, seed, flags, want_all=False):
        """Opt-in mode (flags: 1 = Hartley-normalised solve, 2 = true Sampson distance); flags = 0 is find_fundamental."""
        m = len(matches)
        matches = np.ascontiguousarray(matches, np.int32)
        F, mask = np.zeros((3, 3), np.float32), np.zeros(max(m, 1), np.uint8)
        n, s, b = C.c_int(), C.c_float(), C.c_int()
        Fall, cnt, sc = np.zeros((max(iters, 1), 9), np.float32), np.zeros(max(iters, 1), np.int32), np.zeros(max(iters, 1), np.float32)
        rc = self.lib.vbo_find_fundamental_ex(p1, p2, matches, m, min_items, iters, thr, seed, flags, F, mask, C.byref(n), C.byref(s),
                                              C.byref(b), Fall.ctypes.data, cnt.ctypes.data, sc.ctypes.data)
        out = dict(rc=rc, F=F, mask=mask[:m], n_inliers=n.value, score=np.float32(s.value), best=b.value)
        if want_all:
            out.update(F_all=Fall[:iters], cnt_all=cnt[:iters], score_all=sc[:iters])
        return out

    # --- matcher ---
    def knn2_hamming(self, d1, d2):
        n1 = len(d1)
        idx, dist = np.zeros((n1, 2), np.int32), np.zeros((n1, 2), np.int32)
        self.lib.vbo_knn2_hamming(d1, n1, d2, len(d2), d1.shape[1], idx, dist)
        return idx, dist

    def match_hamming(self, d1, d2, ratio=0.7):
        out = np.zeros((max(len(d1), 1), 2), np.int32)
        m = self.lib.vbo_match_hamming(d1, len(d1), d2, len(d2), d1.shape[1], ratio, out)
        return out[:m].copy()

    def knn2_l2f(self, d1, d2):
        n1 = len(d1)
        idx, dist = np.zeros((n1, 2), np.int32), np.zeros((n1, 2), np.float32)
        self.lib.vbo_knn2_l2f(d1, n1, d2, len(d2), d1.shape[1], idx, dist)
        return idx, dist

    def match_l2f(self, d1, d2, ratio=0.7):
        out = np.zeros((max(len(d1), 1), 2), np.int32)
        m = self.lib.vbo_match_l2f(d1, len(d1), d2, len(d2), d1.shape[1], ratio, out)
        return out[:m].copy()

    # --- downstream of F (src/helpers.cpp) ---
    def essential(self, F, K):
        E = np.zeros((3, 3), np.float32)
        self.lib.vbo_essential(np.ascontiguousarray(F, np.float32), np.ascontiguousarray(K, np.float32), E)
        return E

    def extract_rt(self, F, K):
        R, t = np.zeros((3, 3), np.float32), np.zeros(3, np.float32)
        self.lib.vbo_extract_rt(np.ascontiguousarray(F, np.float32), np.ascontiguousarray(K, np.float32), R, t)
        return R, t

    def triangulate(self, p1, p2, c1, c2):
        p1, p2 = np.ascontiguousarray(p1, np.float32), np.ascontiguousarray(p2, np.float32)
        out = np.zeros((max(len(p1), 1), 4), np.float32)
        self.lib.vbo_triangulate(p1, p2, len(p1), np.ascontiguousarray(c1, np.float32), np.ascontiguousarray(c2, np.float32), out)
        return out[:len(p1)]

    def reprojection_gate(self, points4, c1, c2, ip1, ip2, ids, thr_sq):
        """src/vslam.cpp:186-251 -> (inlier indices, re1 [n], re2 [n], reproj_error)."""
        points4, ip1, ip2 = (np.ascontiguousarray(a, np.float32) for a in (points4, ip1, ip2))
        n = len(points4)
        re1, re2, idx, err = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.int32), C.c_double()
        ids_p = None if ids is None else np.ascontiguousarray(ids, np.int32).ctypes.data
        cnt = self.lib.vbo_reprojection_gate(points4, n, np.ascontiguousarray(c1, np.float32), np.ascontiguousarray(c2, np.float32),
                                             ip1, ip2, ids_p, thr_sq, re1, re2, idx, C.byref(err))
        return idx[:cnt].copy(), re1[:n], re2[:n], err.value

    # --- search by projection (src/vslam.cpp:129-161) ---
    def project_points(self, X, c2):
        X, c2 = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(c2, np.float32)
        out = np.zeros((len(X), 3), np.float32)
        self.lib.vbo_project_points(X, len(X), c2, out)
        return out

    def search_by_projection(self, X, c2, W, H, pts, desc, map_point_ids, obs_off, obs_desc, radius=2.0, dist_thr=64):
        """Returns (assign [n], map_point_ids after, proj_xy [n][2], in_view [n])."""
        X, c2 = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(c2, np.float32)
        pts, desc = np.ascontiguousarray(pts, np.float32), np.ascontiguousarray(desc, np.uint8)
        pre = self.kdtree_build(pts)
        ids = np.ascontiguousarray(map_point_ids, np.int32).copy()
        obs_off = np.ascontiguousarray(obs_off, np.int32)
        obs_desc = np.ascontiguousarray(obs_desc, np.uint8).reshape(-1, desc.shape[1])
        if len(obs_desc) == 0:
            obs_desc = np.zeros((1, desc.shape[1]), np.uint8)
        n = len(X)
        assign, xy, inv = np.zeros(max(n, 1), np.int32), np.zeros((max(n, 1), 2), np.float32), np.zeros(max(n, 1), np.uint8)
        self.lib.vbo_search_by_projection(X, n, c2, W, H, pts, pre, len(pts), desc, desc.shape[1], ids, obs_off, obs_desc,
                                          radius, dist_thr, assign, xy, inv)
        return assign[:n], ids, xy[:n], inv[:n]


    def match_features(self, p1, d1, p2, d2, ratio, min_items, iters, thr, seed):
        out = np.zeros((max(len(d1), 1), 2), np.int32)
        F = np.zeros((3, 3), np.float32)
        nt, bh = C.c_int(), C.c_int()
        r = self.lib.vbo_match_features(p1, d1, len(d1), p2, d2, len(d2), d1.shape[1], ratio, min_items, iters, thr,
                                        seed, out, F, C.byref(nt), C.byref(bh))
        return dict(n=r, matches=out[:max(r, 0)].copy(), F=F, n_tentative=nt.value, best=bh.value)

    # --- kd-tree ---
    def kdtree_build(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        pre = np.zeros(max(len(pts), 1), np.int32)
        self.lib.vbo_kdtree_build(pts, len(pts), pre)
        return pre[:len(pts)]

    def kdtree_nearest(self, pts, pre, q, max_d2=np.inf):
        d2 = C.c_float()
        slot = self.lib.vbo_kdtree_nearest(pts, pre, len(pts), q[0], q[1], max_d2, C.byref(d2))
        return slot, np.float32(d2.value)

    def kdtree_knn(self, pts, pre, q, k, max_d2=np.inf):
        """-> (pre-order slots [found], squared distances [found])."""
        slot, d2 = np.zeros(max(k, 1), np.int32), np.zeros(max(k, 1), np.float32)
        c = self.lib.vbo_kdtree_knn(pts, pre, len(pts), q[0], q[1], k, max_d2, slot, d2)
        return slot[:c].copy(), d2[:c].copy()

    def kdtree_radius(self, pts, pre, q, r, cap=None):
        cap = cap or len(pts)
        out = np.zeros(max(cap, 1), np.int32)
        c = self.lib.vbo_kdtree_radius(pts, pre, len(pts), q[0], q[1], r, out, cap)
        return out[:min(c, cap)].copy(), c


class Ref:
    """The reference's own KDTree.cpp / RansacFilter.cpp (oracle/_ref/libvbref.so), when it was built."""

    @staticmethod
    def path():
        return os.path.join(ORACLE_DIR, "_ref", "libvbref.so")

    @staticmethod
    def available():
        return os.path.exists(Ref.path())

    def __init__(self):
        L = self.lib = C.CDLL(self.path())
        ip = C.POINTER(C.c_int)
        L.vbref_kdtree_build.argtypes = [_f32p, C.c_int, _f32p, ip, ip]
        L.vbref_kdtree_nearest.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_float, _f32p]
        L.vbref_kdtree_radius.restype = C.c_long
        L.vbref_kdtree_radius.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_float, _i32p, _f32p, C.c_long]
        L.vbref_frame_kdtree_build.argtypes = [_f32p, C.c_int, _i64p, ip]
        L.vbref_frame_kdtree_radius.restype = C.c_long
        L.vbref_frame_kdtree_radius.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_float, _i32p, _i64p, C.c_long]
        L.vbref_kdtree_time_ms.restype = C.c_double
        L.vbref_kdtree_time_ms.argtypes = [C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.c_float, C.c_int]
        L.vbref_initialize_sets.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, _i32p]
        L.vbref_find_fundamental.restype = C.c_int
        L.vbref_find_fundamental.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int,
                                             C.c_float, C.c_uint, _f32p, _u8p, ip]
        L.vbref_compute_fundamental.argtypes = [_f32p, _f32p, _f32p]
        L.vbref_residual.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _f32p, C.c_float, _u8p, ip,
                                     C.POINTER(C.c_float)]
        L.vbref_find_fundamental_time_ms.restype = C.c_double
        L.vbref_find_fundamental_time_ms.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, C.c_int,
                                                     C.c_float, C.c_uint, C.c_int]

    def kdtree_build(self, pts):
        n = len(pts)
        pre = np.zeros((max(n, 1), 2), np.float32)
        h, ok = C.c_int(), C.c_int()
        self.lib.vbref_kdtree_build(pts, n, pre, C.byref(h), C.byref(ok))
        return pre[:n], h.value, bool(ok.value)

    def kdtree_nearest(self, pts, q, max_d2=np.inf):
        out = np.zeros((len(q), 2), np.float32)
        self.lib.vbref_kdtree_nearest(pts, len(pts), q, len(q), max_d2, out)
        return out

    def kdtree_radius(self, pts, q, r, cap=1 << 22):
        off = np.zeros(len(q) + 1, np.int32)
        out = np.zeros((cap, 2), np.float32)
        tot = self.lib.vbref_kdtree_radius(pts, len(pts), q, len(q), r, off, out, cap)
        assert tot <= cap
        return off, out[:tot]

    def frame_kdtree_build(self, pts):
        n = len(pts)
        pre = np.zeros(max(n, 1), np.int64)
        h = C.c_int()
        self.lib.vbref_frame_kdtree_build(pts, n, pre, C.byref(h))
        return pre[:n], h.value

    def frame_kdtree_radius(self, pts, q, r, cap=1 << 22):
        off = np.zeros(len(q) + 1, np.int32)
        out = np.zeros(cap, np.int64)
        tot = self.lib.vbref_frame_kdtree_radius(pts, len(pts), q, len(q), r, off, out, cap)
        assert tot <= cap
        return off, out[:tot]

    def initialize_sets(self, n_matches, min_items, iters, seed):
        sets = np.zeros((iters, 8), np.int32)
        self.lib.vbref_initialize_sets(n_matches, min_items, iters, seed, sets)
        return sets

    def find_fundamental(self, p1, p2, matches, min_items, iters, thr, seed):
        m = len(matches)
        F, mask, ml = np.zeros((3, 3), np.float32), np.zeros(max(m, 1), np.uint8), C.c_int()
        ok = self.lib.vbref_find_fundamental(p1, len(p1), p2, len(p2), np.ascontiguousarray(matches, np.int32), m,
                                             min_items, iters, thr, seed, F, mask, C.byref(ml))
        return dict(accepted=bool(ok), F=F, mask=mask[:ml.value])

    def compute_fundamental(self, p1set, p2set):
        F = np.zeros((3, 3), np.float32)
        self.lib.vbref_compute_fundamental(np.ascontiguousarray(p1set, np.float32),
                                           np.ascontiguousarray(p2set, np.float32), F)
        return F

    def residual(self, p1, p2, matches, F, thr):
        m = len(matches)
        mask, n, s = np.zeros(max(m, 1), np.uint8), C.c_int(), C.c_float()
        self.lib.vbref_residual(p1, len(p1), p2, len(p2), np.ascontiguousarray(matches, np.int32), m,
                                np.ascontiguousarray(F, np.float32).reshape(-1), thr, mask, C.byref(n), C.byref(s))
        return mask[:m], n.value, np.float32(s.value)
