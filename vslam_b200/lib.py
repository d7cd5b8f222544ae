"""ctypes binding of libvslam_b200.so (include/vslam_b200.h).

This is plumbing for tests and bench.py; the product is the C ABI itself. There is no fallback of any
kind: if the shared library is missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# VB_LIB_PATH: tools/ may point the binding at an out-of-tree build of the same sources (e.g. a TUNING=1 build with the
# timing-only options compiled in); tests and bench.py always use the in-tree product build.
LIB_PATH = os.environ.get("VB_LIB_PATH") or os.path.join(PKG_DIR, "lib", "libvslam_b200.so")

VB_OK, VB_ERR_INVALID, VB_ERR_CUDA, VB_ERR_TOO_FEW, VB_ERR_CAPACITY, VB_ERR_NO_MODEL = range(6)


class VbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vslam_b200 error {code}: {msg}")
        self.code = code


class PairParams(C.Structure):
    _fields_ = [("ratio", C.c_double), ("min_items", C.c_int32), ("max_iterations", C.c_uint32),
                ("threshold", C.c_float), ("seed0", C.c_uint32)]


class PairResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_tentative", C.c_int32), ("n_matches", C.c_int32),
                ("best_hyp", C.c_int32), ("n_inliers", C.c_int32), ("score", C.c_float), ("F", C.c_float * 9)]


PAIR_RESULT_DTYPE = np.dtype([("status", "<i4"), ("n_tentative", "<i4"), ("n_matches", "<i4"), ("best_hyp", "<i4"),
                              ("n_inliers", "<i4"), ("score", "<f4"), ("F", "<f4", (9,))])
assert PAIR_RESULT_DTYPE.itemsize == C.sizeof(PairResult) == 60

EXPORTS = [
    "vb_version", "vb_last_error", "vb_create", "vb_destroy", "vb_set_stream", "vb_set_option", "vb_reset_options", "vb_synchronize", "vb_launch_count",
    "vb_kdtree_build", "vb_kdtree_build_d", "vb_kdtree_build_batch_d", "vb_kdtree_free_batch", "vb_kdtree_import", "vb_kdtree_free", "vb_kdtree_size", "vb_kdtree_height", "vb_kdtree_export",
    "vb_kdtree_nearest", "vb_kdtree_nearest_d", "vb_kdtree_knn", "vb_kdtree_knn_d", "vb_kdtree_radius", "vb_kdtree_radius_d",
    "vb_knn2_hamming", "vb_match_hamming", "vb_knn2_l2f", "vb_match_l2f",
    "vb_ransac_fundamental", "vb_ransac_fundamental_ex", "vb_ransac_hypotheses", "vb_ransac_score", "vb_ransac_score_d", "vb_ransac_counts", "vb_ransac_counts_d", "vb_ransac_prune_stats", "vb_ransac_solve8", "vb_ransac_sample_sets", "vb_ransac_residual",
    "vb_match_features", "vb_match_features_l2f", "vb_match_features_l2f_d", "vb_pairs_run", "vb_pairs_run_d", "vb_pairs_submit", "vb_pairs_submit_d", "vb_pairs_wait", "vb_pairs_run_compact",
    "vb_host_alloc", "vb_host_free", "vb_host_register", "vb_host_unregister",
    "vb_multi_create", "vb_multi_destroy", "vb_multi_device_count", "vb_multi_pairs_submit", "vb_multi_pairs_wait", "vb_multi_pairs_run", "vb_search_by_projection", "vb_extract_rt", "vb_triangulate", "vb_triangulate_gated", "vb_profile_enable", "vb_profile_last_ms", "vb_probe_tensor_peak",
]


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvslam_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() / make -C vslam_b200/csrc "
                                "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, u64, f32, f64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_float, C.c_double
    L.vb_version.restype = C.c_int
    L.vb_last_error.restype = C.c_char_p
    L.vb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.vb_destroy.argtypes = [vp]
    L.vb_set_stream.argtypes = [vp, vp]
    L.vb_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.vb_reset_options.argtypes = [vp]
    L.vb_synchronize.argtypes = [vp]
    L.vb_launch_count.restype = u64
    L.vb_launch_count.argtypes = [vp]
    L.vb_kdtree_build.argtypes = [vp, vp, u32, C.POINTER(vp)]
    L.vb_kdtree_build_d.argtypes = [vp, vp, u32, C.POINTER(vp)]
    L.vb_kdtree_build_batch_d.argtypes = [vp, vp, u32, u32, C.POINTER(vp)]
    L.vb_kdtree_free_batch.argtypes = [C.POINTER(vp), u32]
    L.vb_kdtree_import.argtypes = [vp, vp, vp, u32, C.POINTER(vp)]
    L.vb_kdtree_free.argtypes = [vp]
    L.vb_kdtree_size.restype = u32
    L.vb_kdtree_size.argtypes = [vp]
    L.vb_kdtree_height.restype = u32
    L.vb_kdtree_height.argtypes = [vp]
    L.vb_kdtree_export.argtypes = [vp, vp, vp]
    L.vb_kdtree_nearest.argtypes = [vp, vp, u32, f32, vp, vp, vp]
    L.vb_kdtree_nearest_d.argtypes = [vp, vp, u32, f32, vp, vp, vp]
    L.vb_kdtree_knn.argtypes = [vp, vp, u32, u32, f32, vp, vp, vp]
    L.vb_kdtree_knn_d.argtypes = [vp, vp, u32, u32, f32, vp, vp, vp]
    L.vb_kdtree_radius.argtypes = [vp, vp, u32, f32, vp, vp, u64, C.POINTER(u64)]
    L.vb_kdtree_radius_d.argtypes = [vp, vp, u32, f32, vp, vp, u64, C.POINTER(u64)]
    L.vb_knn2_hamming.argtypes = [vp, vp, u32, vp, u32, u32, vp, vp]
    L.vb_match_hamming.argtypes = [vp, vp, u32, vp, u32, u32, f64, vp, C.POINTER(u32)]
    L.vb_knn2_l2f.argtypes = [vp, vp, u32, vp, u32, u32, vp, vp]
    L.vb_match_l2f.argtypes = [vp, vp, u32, vp, u32, u32, f64, vp, C.POINTER(u32)]
    L.vb_ransac_fundamental.argtypes = [vp, vp, u32, vp, u32, vp, u32, C.c_int, u32, f32, u32, vp, vp,
                                        C.POINTER(i32), C.POINTER(f32), C.POINTER(i32)]
    L.vb_ransac_fundamental_ex.argtypes = [vp, vp, u32, vp, u32, vp, u32, C.c_int, u32, f32, u32, u32, vp, vp,
                                           C.POINTER(i32), C.POINTER(f32), C.POINTER(i32)]
    L.vb_ransac_hypotheses.argtypes = [vp, vp, u32, vp, u32, vp, u32, C.c_int, u32, f32, u32, vp, vp, vp, vp]
    L.vb_ransac_score.argtypes = [vp, vp, u32, vp, u32, f32, vp, vp]
    L.vb_ransac_score_d.argtypes = [vp, vp, u32, vp, u32, f32, vp, vp]
    L.vb_ransac_counts.argtypes = [vp, vp, u32, vp, u32, f32, vp]
    L.vb_ransac_counts_d.argtypes = [vp, vp, u32, vp, u32, f32, vp]
    L.vb_ransac_prune_stats.argtypes = [vp, vp, vp, C.c_int]
    L.vb_ransac_solve8.argtypes = [vp, vp, vp, u32, vp]
    L.vb_ransac_sample_sets.argtypes = [vp, u32, C.c_int, u32, u32, vp]
    L.vb_ransac_residual.argtypes = [vp, vp, u32, vp, u32, vp, u32, vp, f32, vp, C.POINTER(i32), C.POINTER(f32)]
    L.vb_match_features.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, C.POINTER(PairParams), vp, C.POINTER(PairResult)]
    L.vb_match_features_l2f.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, C.POINTER(PairParams), vp, C.POINTER(PairResult)]
    L.vb_match_features_l2f_d.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, C.POINTER(PairParams), vp, vp]
    L.vb_pairs_run.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp]
    L.vb_pairs_run_d.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp]
    L.vb_pairs_submit.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp, vp, u64, C.POINTER(C.c_int)]
    L.vb_pairs_submit_d.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp, C.POINTER(C.c_int)]
    L.vb_pairs_wait.argtypes = [vp, C.c_int, C.POINTER(u64)]
    L.vb_pairs_run_compact.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp, vp, u64, C.POINTER(u64)]
    L.vb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.vb_host_free.argtypes = [vp]
    L.vb_host_register.argtypes = [vp, C.c_size_t]
    L.vb_host_unregister.argtypes = [vp]
    L.vb_multi_create.argtypes = [C.POINTER(C.c_int), u32, C.POINTER(vp)]
    L.vb_multi_destroy.argtypes = [vp]
    L.vb_multi_device_count.restype = u32
    L.vb_multi_device_count.argtypes = [vp]
    L.vb_multi_pairs_submit.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp, vp, u64, C.POINTER(C.c_int)]
    L.vb_multi_pairs_wait.argtypes = [vp, C.c_int, C.POINTER(u64)]
    L.vb_multi_pairs_run.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(PairParams), vp, vp, vp, u64, C.POINTER(u64)]
    L.vb_search_by_projection.argtypes = [vp, vp, vp, u32, vp, C.c_int, C.c_int, vp, u32, vp, vp, vp, f32, u32, vp, vp, vp,
                                          C.POINTER(u32)]
    L.vb_extract_rt.argtypes = [vp, vp, u32, vp, vp, vp, vp]
    L.vb_triangulate.argtypes = [vp, vp, vp, u32, vp, vp, vp]
    L.vb_probe_tensor_peak.argtypes = [vp, C.c_int, u32, u32, u32, C.POINTER(f32), C.POINTER(f64)]
    L.vb_triangulate_gated.argtypes = [vp, vp, vp, u32, vp, vp, vp, f32, vp, vp, vp, vp, C.POINTER(u32), C.POINTER(f64)]
    L.vb_profile_enable.argtypes = [vp, C.c_int]
    L.vb_profile_last_ms.restype = f32
    L.vb_profile_last_ms.argtypes = [vp, C.c_char_p]
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


class Context:
    """One vb_ctx. Host-array methods mirror the C entry points one to one."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.vb_create(device, C.byref(h))
        if rc != VB_OK:
            raise VbError(rc, self.L.vb_last_error().decode())
        self.h = h
        self.device = device
        # tools/ convenience (Python plumbing only — the library itself never reads the environment):
        # VB_OPTIONS="hamming_tc=1,prune_rounds=6" applies vb_set_option to every context this process creates
        for kv in filter(None, os.environ.get("VB_OPTIONS", "").split(",")):
            name, _, val = kv.partition("=")
            self.set_option(name.strip(), int(val))

    def close(self):
        if getattr(self, "h", None):
            self.L.vb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, ok=(VB_OK,)):
        if rc not in ok:
            raise VbError(rc, self.L.vb_last_error().decode())
        return rc

    def set_stream(self, cuda_stream_ptr: int | None):
        self._chk(self.L.vb_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_option(self, name: str, value: int):
        """Force one of the equivalent code paths (include/vslam_b200.h, vb_set_option)."""
        self._chk(self.L.vb_set_option(self.h, name.encode(), int(value)))

    def reset_options(self):
        self._chk(self.L.vb_reset_options(self.h))

    def synchronize(self):
        self._chk(self.L.vb_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.L.vb_launch_count(self.h))

    def probe_tensor_peak(self, kind: int, n_cols: int = 256, iters: int = 4096, reps: int = 5) -> float:
        """TFLOP/s of back-to-back tcgen05.mma of one kind (0 mxf4, 1 f8f6f4, 2 f16/bf16), nothing else running."""
        ms, fl = C.c_float(), C.c_double()
        self._chk(self.L.vb_probe_tensor_peak(self.h, kind, n_cols, iters, reps, C.byref(ms), C.byref(fl)))
        return fl.value / (ms.value * 1e-3) / 1e12

    def profile(self, on: bool):
        self._chk(self.L.vb_profile_enable(self.h, int(on)))

    def profile_ms(self, name: str) -> float:
        return float(self.L.vb_profile_last_ms(self.h, name.encode()))

    # ---- kd-tree ----
    def kdtree_build(self, pts):
        pts = _f32(pts)
        t = C.c_void_p()
        self._chk(self.L.vb_kdtree_build(self.h, _ptr(pts), len(pts), C.byref(t)))
        return KDTreeHandle(self, t, len(pts))

    # ---- matcher ----
    def knn2_hamming(self, d1, d2):
        d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
        idx, dist = np.zeros((len(d1), 2), np.int32), np.zeros((len(d1), 2), np.int32)
        self._chk(self.L.vb_knn2_hamming(self.h, _ptr(d1), len(d1), _ptr(d2), len(d2), d1.shape[1], _ptr(idx), _ptr(dist)))
        return idx, dist

    def match_hamming(self, d1, d2, ratio=0.7):
        d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
        out, m = np.zeros((max(len(d1), 1), 2), np.int32), C.c_uint32()
        self._chk(self.L.vb_match_hamming(self.h, _ptr(d1), len(d1), _ptr(d2), len(d2), d1.shape[1], ratio, _ptr(out), C.byref(m)))
        return out[:m.value].copy()

    def knn2_l2f(self, d1, d2):
        d1, d2 = _f32(d1), _f32(d2)
        idx, dist = np.zeros((len(d1), 2), np.int32), np.zeros((len(d1), 2), np.float32)
        self._chk(self.L.vb_knn2_l2f(self.h, _ptr(d1), len(d1), _ptr(d2), len(d2), d1.shape[1], _ptr(idx), _ptr(dist)))
        return idx, dist

    def match_l2f(self, d1, d2, ratio=0.7):
        d1, d2 = _f32(d1), _f32(d2)
        out, m = np.zeros((max(len(d1), 1), 2), np.int32), C.c_uint32()
        self._chk(self.L.vb_match_l2f(self.h, _ptr(d1), len(d1), _ptr(d2), len(d2), d1.shape[1], ratio, _ptr(out), C.byref(m)))
        return out[:m.value].copy()

    # ---- search by projection (reference src/vslam.cpp:129-161) ----
    def search_by_projection(self, tree, X, c2, W, H, desc, map_point_ids, obs_off, obs_desc, radius=2.0, dist_thr=64):
        """Returns (assign [n], map_point_ids after, proj_xy [n][2], in_view [n], n_claimed)."""
        X, c2 = _f32(X), _f32(c2)
        desc = np.ascontiguousarray(desc, np.uint8)
        ids = np.ascontiguousarray(map_point_ids, np.int32).copy()
        obs_off = np.ascontiguousarray(obs_off, np.uint32)
        obs_desc = np.ascontiguousarray(obs_desc, np.uint8)
        n = len(X)
        assign = np.zeros(max(n, 1), np.int32)
        xy, inv, cnt = np.zeros((max(n, 1), 2), np.float32), np.zeros(max(n, 1), np.uint8), C.c_uint32()
        self._chk(self.L.vb_search_by_projection(self.h, tree.h, _ptr(X), n, _ptr(c2), int(W), int(H), _ptr(desc),
                                                 desc.shape[1], _ptr(ids), _ptr(obs_off), _ptr(obs_desc), float(radius),
                                                 int(dist_thr), _ptr(assign), _ptr(xy), _ptr(inv), C.byref(cnt)))
        return assign[:n], ids, xy[:n], inv[:n], cnt.value

    # ---- downstream of F (reference src/helpers.cpp) ----
    def extract_rt(self, F, K):
        """F [P][3][3] (or [3][3]) -> (R [P][3][3], t [P][3], E [P][3][3])."""
        F = _f32(F).reshape(-1, 9)
        K = _f32(K).reshape(9)
        P = len(F)
        R, t, E = np.zeros((max(P, 1), 9), np.float32), np.zeros((max(P, 1), 3), np.float32), np.zeros((max(P, 1), 9), np.float32)
        self._chk(self.L.vb_extract_rt(self.h, _ptr(F), P, _ptr(K), _ptr(R), _ptr(t), _ptr(E)))
        return R[:P].reshape(-1, 3, 3), t[:P], E[:P].reshape(-1, 3, 3)

    def triangulate(self, p1, p2, c1, c2):
        p1, p2, c1, c2 = _f32(p1), _f32(p2), _f32(c1).reshape(12), _f32(c2).reshape(12)
        out = np.zeros((max(len(p1), 1), 4), np.float32)
        self._chk(self.L.vb_triangulate(self.h, _ptr(p1), _ptr(p2), len(p1), _ptr(c1), _ptr(c2), _ptr(out)))
        return out[:len(p1)]

    def triangulate_gated(self, p1, p2, c1, c2, ids, thr_sq=4.0):
        """-> (points4 [n][4], inlier indices, re1 [n], re2 [n], reproj_error)."""
        p1, p2, c1, c2 = _f32(p1), _f32(p2), _f32(c1).reshape(12), _f32(c2).reshape(12)
        n = len(p1)
        ids = None if ids is None else np.ascontiguousarray(ids, np.int32)
        out = np.zeros((max(n, 1), 4), np.float32)
        re1, re2, idx = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.uint32)
        cnt, err = C.c_uint32(), C.c_double()
        self._chk(self.L.vb_triangulate_gated(self.h, _ptr(p1), _ptr(p2), n, _ptr(c1), _ptr(c2), _ptr(ids), thr_sq, _ptr(out),
                                              _ptr(re1), _ptr(re2), _ptr(idx), C.byref(cnt), C.byref(err)))
        return out[:n], idx[:cnt.value].astype(np.int32), re1[:n], re2[:n], err.value

    # ---- ransac ----
    def ransac_fundamental(self, p1, p2, matches, min_items=8, iters=100, thr=10.0, seed=0, flags=0):
        p1, p2, matches = _f32(p1), _f32(p2), np.ascontiguousarray(matches, np.int32)
        m = len(matches)
        F, mask = np.zeros((3, 3), np.float32), np.zeros(max(m, 1), np.uint8)
        n, s, b = C.c_int32(), C.c_float(), C.c_int32()
        if flags:
            rc = self.L.vb_ransac_fundamental_ex(self.h, _ptr(p1), len(p1), _ptr(p2), len(p2), _ptr(matches), m, min_items, iters,
                                                 thr, seed, flags, _ptr(F), _ptr(mask), C.byref(n), C.byref(s), C.byref(b))
        else:
            rc = self.L.vb_ransac_fundamental(self.h, _ptr(p1), len(p1), _ptr(p2), len(p2), _ptr(matches), m, min_items, iters,
                                              thr, seed, _ptr(F), _ptr(mask), C.byref(n), C.byref(s), C.byref(b))
        self._chk(rc, ok=(VB_OK, VB_ERR_TOO_FEW, VB_ERR_NO_MODEL))
        return dict(rc=rc, F=F, mask=mask[:m], n_inliers=n.value, score=np.float32(s.value), best=b.value)

    def ransac_hypotheses(self, p1, p2, matches, min_items=8, iters=100, thr=10.0, seed=0):
        p1, p2, matches = _f32(p1), _f32(p2), np.ascontiguousarray(matches, np.int32)
        sets, Fall = np.zeros((iters, 8), np.int32), np.zeros((iters, 9), np.float32)
        cnt, sc = np.zeros(iters, np.int32), np.zeros(iters, np.float32)
        self._chk(self.L.vb_ransac_hypotheses(self.h, _ptr(p1), len(p1), _ptr(p2), len(p2), _ptr(matches), len(matches),
                                              min_items, iters, thr, seed, _ptr(sets), _ptr(Fall), _ptr(cnt), _ptr(sc)))
        return dict(sets=sets, F_all=Fall, cnt_all=cnt, score_all=sc)

    def ransac_score(self, corr, F, thr=10.0):
        corr, F = _f32(corr), _f32(F).reshape(-1, 9)
        cnt, sc = np.zeros(len(F), np.int32), np.zeros(len(F), np.float32)
        self._chk(self.L.vb_ransac_score(self.h, _ptr(corr), len(corr), _ptr(F), len(F), thr, _ptr(cnt), _ptr(sc)))
        return cnt, sc

    def ransac_counts(self, corr, F, thr=10.0):
        corr, F = _f32(corr), _f32(F).reshape(-1, 9)
        cnt = np.zeros(len(F), np.int32)
        self._chk(self.L.vb_ransac_counts(self.h, _ptr(corr), len(corr), _ptr(F), len(F), thr, _ptr(cnt)))
        return cnt

    def ransac_prune_stats(self, reset=False):
        """(evaluations performed, evaluations a full count would take) of the pipeline's bounded counting so far."""
        ev, tot = C.c_uint64(0), C.c_uint64(0)
        self._chk(self.L.vb_ransac_prune_stats(self.h, C.byref(ev), C.byref(tot), 1 if reset else 0))
        return ev.value, tot.value

    def ransac_solve8(self, p1set, p2set):
        p1set, p2set = _f32(p1set).reshape(-1, 8, 2), _f32(p2set).reshape(-1, 8, 2)
        F = np.zeros((len(p1set), 3, 3), np.float32)
        self._chk(self.L.vb_ransac_solve8(self.h, _ptr(p1set), _ptr(p2set), len(p1set), _ptr(F)))
        return F

    # ---- whole pair(s) ----
    @staticmethod
    def params(ratio=0.7, min_items=8, iters=100, thr=10.0, seed0=0):
        return PairParams(ratio, min_items, iters, thr, seed0)

    def match_features(self, p1, d1, p2, d2, prm: PairParams):
        p1, p2 = _f32(p1), _f32(p2)
        d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
        out, res = np.zeros((max(len(d1), 1), 2), np.int32), PairResult()
        self._chk(self.L.vb_match_features(self.h, _ptr(p1), _ptr(d1), len(d1), _ptr(p2), _ptr(d2), len(d2), d1.shape[1],
                                           C.byref(prm), _ptr(out), C.byref(res)))
        return dict(status=res.status, n=res.n_matches, matches=out[:max(res.n_matches, 0)].copy(),
                    F=np.array(res.F, np.float32).reshape(3, 3), n_tentative=res.n_tentative, best=res.best_hyp,
                    n_inliers=res.n_inliers, score=np.float32(res.score))

    def match_features_l2f(self, p1, d1, p2, d2, prm: PairParams):
        p1, p2, d1, d2 = _f32(p1), _f32(p2), _f32(d1), _f32(d2)
        out, res = np.zeros((max(len(d1), 1), 2), np.int32), PairResult()
        self._chk(self.L.vb_match_features_l2f(self.h, _ptr(p1), _ptr(d1), len(d1), _ptr(p2), _ptr(d2), len(d2), d1.shape[1],
                                               C.byref(prm), _ptr(out), C.byref(res)))
        return dict(status=res.status, n=res.n_matches, matches=out[:max(res.n_matches, 0)].copy(),
                    F=np.array(res.F, np.float32).reshape(3, 3), n_tentative=res.n_tentative, best=res.best_hyp,
                    n_inliers=res.n_inliers, score=np.float32(res.score))

    def pairs_run(self, pts, desc, prm: PairParams, want_matches=True):
        pts, desc = _f32(pts), np.ascontiguousarray(desc, np.uint8)
        nframes, k = pts.shape[0], pts.shape[1]
        res = np.zeros(max(nframes - 1, 1), PAIR_RESULT_DTYPE)
        out = np.zeros((max(nframes - 1, 1), k, 2), np.int32) if want_matches else None
        self._chk(self.L.vb_pairs_run(self.h, _ptr(pts), _ptr(desc), nframes, k, desc.shape[2], C.byref(prm), _ptr(res), _ptr(out)))
        return res[:nframes - 1], (out[:nframes - 1] if want_matches else None)

    def pairs_submit(self, pts, desc, prm: PairParams, res, offsets, matches16):
        """Asynchronous: the arrays must stay alive (and untouched) until pairs_wait(ticket)."""
        nframes, k = pts.shape[0], pts.shape[1]
        t = C.c_int(-1)
        self._chk(self.L.vb_pairs_submit(self.h, _ptr(pts), _ptr(desc), nframes, k, desc.shape[2], C.byref(prm), _ptr(res),
                                         _ptr(offsets), _ptr(matches16), 0 if matches16 is None else len(matches16), C.byref(t)))
        return t.value

    def pairs_wait(self, ticket):
        tot = C.c_uint64(0)
        self._chk(self.L.vb_pairs_wait(self.h, ticket, C.byref(tot)))
        return tot.value

    def pairs_run_compact(self, pts, desc, prm: PairParams, cap=None):
        pts, desc = _f32(pts), np.ascontiguousarray(desc, np.uint8)
        P, k = pts.shape[0] - 1, pts.shape[1]
        res = np.zeros(max(P, 1), PAIR_RESULT_DTYPE)
        off = np.zeros(max(P, 1), np.uint32)
        m16 = np.zeros((max(P * k if cap is None else cap, 1), 2), np.uint16)
        self.pairs_wait(self.pairs_submit(pts, desc, prm, res, off, m16))
        return res[:P], off[:P], m16


def pinned_empty(L, shape, dtype):
    """numpy array over vb_host_alloc'd (pinned, portable) memory; keep the returned array alive, free with pinned_free."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = L.vb_host_alloc(max(n, 1), C.byref(p))
    if rc != VB_OK:
        raise VbError(rc, L.vb_last_error().decode())
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)
    return a, p


def unpack_compact(res, offsets, matches16):
    """[(n_i, 2) int32 arrays] from the compact download."""
    return [matches16[int(o):int(o) + int(n)].astype(np.int32) for o, n in zip(offsets, res["n_matches"])]


class Multi:
    """vb_multi: one host thread + context per listed GPU inside this process."""

    def __init__(self, devices):
        self.L = load_library()
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.L.vb_multi_create(arr, len(devices), C.byref(h))
        if rc != VB_OK:
            raise VbError(rc, self.L.vb_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.vb_multi_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != VB_OK:
            raise VbError(rc, self.L.vb_last_error().decode())

    def submit(self, pts, desc, prm, res, offsets, matches16):
        nframes, k = pts.shape[0], pts.shape[1]
        t = C.c_int(-1)
        self._chk(self.L.vb_multi_pairs_submit(self.h, _ptr(pts), _ptr(desc), nframes, k, desc.shape[2], C.byref(prm), _ptr(res),
                                               _ptr(offsets), _ptr(matches16), 0 if matches16 is None else len(matches16),
                                               C.byref(t)))
        return t.value

    def wait(self, ticket):
        tot = C.c_uint64(0)
        self._chk(self.L.vb_multi_pairs_wait(self.h, ticket, C.byref(tot)))
        return tot.value

    def pairs_run(self, pts, desc, prm):
        pts, desc = _f32(pts), np.ascontiguousarray(desc, np.uint8)
        P, k = pts.shape[0] - 1, pts.shape[1]
        res = np.zeros(max(P, 1), PAIR_RESULT_DTYPE)
        off, m16 = np.zeros(max(P, 1), np.uint32), np.zeros((max(P * k, 1), 2), np.uint16)
        self.wait(self.submit(pts, desc, prm, res, off, m16))
        return res[:P], off[:P], m16


class KDTreeHandle:
    def __init__(self, ctx: "Context", handle, n):
        self.ctx, self.h, self.n = ctx, handle, n

    def free(self):
        if self.h:
            self.ctx.L.vb_kdtree_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def height(self):
        return int(self.ctx.L.vb_kdtree_height(self.h))

    def export(self):
        idx, pts = np.zeros(max(self.n, 1), np.uint32), np.zeros((max(self.n, 1), 2), np.float32)
        self.ctx._chk(self.ctx.L.vb_kdtree_export(self.h, _ptr(idx), _ptr(pts)))
        return idx[:self.n], pts[:self.n]

    def nearest(self, q, max_d2=np.inf):
        q = _f32(q).reshape(-1, 2)
        pt, idx, d2 = np.zeros((len(q), 2), np.float32), np.zeros(len(q), np.int32), np.zeros(len(q), np.float32)
        self.ctx._chk(self.ctx.L.vb_kdtree_nearest(self.h, _ptr(q), len(q), max_d2, _ptr(pt), _ptr(idx), _ptr(d2)))
        return pt, idx, d2

    def knn(self, q, k, max_d2=np.inf):
        """-> (idx [nq][k] original indices or -1, d2 [nq][k], count [nq])."""
        q = _f32(q).reshape(-1, 2)
        idx, d2, cnt = np.zeros((len(q), k), np.int32), np.zeros((len(q), k), np.float32), np.zeros(len(q), np.uint32)
        self.ctx._chk(self.ctx.L.vb_kdtree_knn(self.h, _ptr(q), len(q), k, max_d2, _ptr(idx), _ptr(d2), _ptr(cnt)))
        return idx, d2, cnt

    def radius(self, q, r, cap=None):
        q = _f32(q).reshape(-1, 2)
        cap = cap if cap is not None else max(16 * len(q), 1024)
        while True:
            off, out, tot = np.zeros(len(q) + 1, np.uint32), np.zeros(max(cap, 1), np.uint32), C.c_uint64()
            rc = self.ctx.L.vb_kdtree_radius(self.h, _ptr(q), len(q), r, _ptr(off), _ptr(out), cap, C.byref(tot))
            if rc == VB_ERR_CAPACITY and tot.value > cap:
                cap = int(tot.value)
                continue
            self.ctx._chk(rc)
            return off, out[:tot.value].copy()
