"""ctypes binding of oracle/libcvoracle.so — the CPU restatement of cv::BFMatcher::knnMatch and
cv::findHomography used as the parity checker (reference call sites: src/TestsDetector.cpp:59-84).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Parity is pinned to cv2 4.13.0 via tests/golden/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcvoracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("cvoracle.c", "cvoracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libcvoracle.so"])
    return _SO


class RansacInfo(C.Structure):
    _fields_ = [("iters_run", C.c_int), ("niters_final", C.c_int), ("best_iter", C.c_int),
                ("best_count", C.c_int), ("draws", C.c_int64), ("lm_iters", C.c_int),
                ("n_refit", C.c_int)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.cvo_l2sqr.restype = C.c_float
        _lib.cvo_invert_eig.restype = C.c_double
        _lib.cvo_det3.restype = C.c_double
        _lib.cvo_rng_next.restype = C.c_uint32
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def knn2(q, t, nthreads=1):
    """cv::BFMatcher(NORM_L2).knnMatch(q, t, k=2) -> idx [nq,2] int32 (-1 absent), dist [nq,2] f32."""
    q = _f32(q); t = _f32(t)
    nq = q.shape[0]; nt = t.shape[0]
    dim = q.shape[1] if q.ndim == 2 else (t.shape[1] if t.ndim == 2 else 128)
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.zeros((nq, 2), np.float32)
    if nq:
        lib().cvo_knn2(_p(q, C.c_float), nq, _p(t, C.c_float), nt, dim,
                       _p(idx, C.c_int32), _p(dist, C.c_float), int(nthreads))
    return idx, dist


def ratio(idx, dist, ratio_thr=0.9):
    idx = np.ascontiguousarray(idx, np.int32); dist = _f32(dist)
    acc = np.zeros(idx.shape[0], np.uint8)
    lib().cvo_ratio(_p(idx, C.c_int32), _p(dist, C.c_float), idx.shape[0], C.c_float(ratio_thr),
                    _p(acc, C.c_uint8))
    return acc


def l2sqr(a, b):
    a = _f32(a); b = _f32(b)
    return float(lib().cvo_l2sqr(_p(a, C.c_float), _p(b, C.c_float), a.shape[0]))


def jacobi(A):
    A = np.array(A, np.float64, order="C")
    n = A.shape[0]
    W = np.zeros(n); V = np.zeros((n, n))
    lib().cvo_jacobi(_p(A, C.c_double), _p(W, C.c_double), _p(V, C.c_double), n)
    return W, V


def solve_eig(A, b):
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64).ravel()
    x = np.zeros(A.shape[0])
    lib().cvo_solve_eig(_p(A, C.c_double), _p(b, C.c_double), _p(x, C.c_double), A.shape[0])
    return x


def invert_eig(A):
    A = np.ascontiguousarray(A, np.float64)
    Ai = np.zeros_like(A)
    lib().cvo_invert_eig(_p(A, C.c_double), _p(Ai, C.c_double), A.shape[0])
    return Ai


def run_kernel(src, dst):
    src = _f32(src).reshape(-1, 2); dst = _f32(dst).reshape(-1, 2)
    H = np.zeros(9)
    ok = lib().cvo_run_kernel(_p(src, C.c_float), _p(dst, C.c_float), src.shape[0], _p(H, C.c_double))
    return (H.reshape(3, 3) if ok else None)


def check_subset(ms1, ms2):
    ms1 = _f32(ms1).reshape(-1, 2); ms2 = _f32(ms2).reshape(-1, 2)
    return bool(lib().cvo_check_subset(_p(ms1, C.c_float), _p(ms2, C.c_float), ms1.shape[0]))


def compute_error(src, dst, H):
    src = _f32(src).reshape(-1, 2); dst = _f32(dst).reshape(-1, 2)
    H = np.ascontiguousarray(H, np.float64).ravel()
    err = np.zeros(src.shape[0], np.float32)
    lib().cvo_compute_error(_p(src, C.c_float), _p(dst, C.c_float), src.shape[0], _p(H, C.c_double),
                            _p(err, C.c_float))
    return err


def update_num_iters(p, ep, model_points, max_iters):
    return int(lib().cvo_update_num_iters(C.c_double(p), C.c_double(ep), model_points, max_iters))


def lm_refine(src, dst, H, max_iters=10):
    src = _f32(src).reshape(-1, 2); dst = _f32(dst).reshape(-1, 2)
    H = np.array(H, np.float64).ravel().copy()
    it = lib().cvo_lm_refine(_p(src, C.c_float), _p(dst, C.c_float), src.shape[0], _p(H, C.c_double),
                             max_iters)
    return H.reshape(3, 3), it


def lm_refine_affine(src, dst, A, max_iters=10):
    """LM refine step of cv2.estimateAffine2D (test handle for pinning the LM solver)."""
    src = _f32(src).reshape(-1, 2); dst = _f32(dst).reshape(-1, 2)
    A = np.array(A, np.float64).ravel().copy()
    it = lib().cvo_lm_refine_affine(_p(src, C.c_float), _p(dst, C.c_float), src.shape[0],
                                    _p(A, C.c_double), max_iters)
    return A.reshape(2, 3), it


def ransac_stage(src, dst, thr=5.0, conf=0.995, max_iters=2000, want_trace=False):
    src = _f32(src).reshape(-1, 2); dst = _f32(dst).reshape(-1, 2)
    n = src.shape[0]
    H = np.zeros(9); mask = np.zeros(max(n, 1), np.uint8)
    info = RansacInfo()
    samples = np.full((max(max_iters, 1), 4), -1, np.int32) if want_trace else None
    counts = np.full(max(max_iters, 1), -1, np.int32) if want_trace else None
    ok = lib().cvo_ransac_stage(_p(src, C.c_float), _p(dst, C.c_float), n, C.c_double(thr),
                                C.c_double(conf), max_iters, _p(H, C.c_double), _p(mask, C.c_uint8),
                                _p(samples, C.c_int32) if want_trace else None,
                                _p(counts, C.c_int32) if want_trace else None, C.byref(info))
    out = {"found": bool(ok), "H": H.reshape(3, 3), "mask": mask[:n], "info": info.asdict()}
    if want_trace:
        out["samples"] = samples; out["counts"] = counts
    return out


def find_homography(src, dst, method=8, thr=5.0, max_iters=2000, conf=0.995):
    """cv2.findHomography(src, dst, method, thr, maxIters=, confidence=) -> dict(found,H,mask,...)."""
    src = _f32(src).reshape(-1, 2); dst = _f32(dst).reshape(-1, 2)
    n = src.shape[0]
    H = np.zeros(9); mask = np.zeros(max(n, 1), np.uint8); rmask = np.zeros(max(n, 1), np.uint8)
    found = C.c_int(0); info = RansacInfo()
    rc = lib().cvo_find_homography(_p(src, C.c_float), _p(dst, C.c_float), n, method, C.c_double(thr),
                                   max_iters, C.c_double(conf), _p(H, C.c_double), _p(mask, C.c_uint8),
                                   C.byref(found), _p(rmask, C.c_uint8), C.byref(info))
    if rc != 0:
        raise ValueError("findHomography needs at least 4 correspondences")
    return {"found": bool(found.value), "H": H.reshape(3, 3), "mask": mask[:n],
            "ransac_mask": rmask[:n], "info": info.asdict()}


def det3(H):
    H = np.ascontiguousarray(H, np.float64).ravel()
    return float(lib().cvo_det3(_p(H, C.c_double)))


ACCEPT, LT4_MATCHES, H_EMPTY, LT4_INLIERS, DET_REJECT = range(5)


def gate(n_good, found, mask, H, min_inliers=4, det_lo=0.1, det_hi=10.0):
    mask = np.ascontiguousarray(mask, np.uint8)
    H = np.ascontiguousarray(H, np.float64).ravel()
    return int(lib().cvo_gate(int(n_good), int(found), _p(mask, C.c_uint8), _p(H, C.c_double),
                              min_inliers, C.c_float(det_lo), C.c_float(det_hi)))
