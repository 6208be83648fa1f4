"""Host build of csrc/homography_math.cuh (the scalar routines the CUDA verify-stage kernels call)
against the oracle and the cv2 goldens — catches logic errors on the CPU before a GPU run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostmath", "hostmath.cpp")
SO = os.path.join(HERE, "hostmath", "libhostmath.so")
CSRC = os.path.join(HERE, "..", "computervision_objectdetection_featurematching_b200", "csrc")
HDR = os.path.join(CSRC, "homography_math.cuh")
HDR2 = os.path.join(CSRC, "jacobi_thread.cuh")


@pytest.fixture(scope="module")
def hm():
    if (not os.path.exists(SO)) or max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2)) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.hm_reproj_err.restype = C.c_float
    lib.hm_reproj_err.argtypes = [C.c_void_p] + [C.c_float] * 4
    lib.hm_det3.restype = C.c_double
    lib.hm_update_num_iters.argtypes = [C.c_double, C.c_double, C.c_int]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_run_kernel_and_lm_match_cv2_goldens(hm, gsynth):
    for s, d, H in zip(gsynth["dlt_src"], gsynth["dlt_dst"], gsynth["dlt_H"]):
        out = np.zeros(9)
        assert hm.hm_run_kernel(_p(s), _p(d), 4, _p(out)) == 1
        assert np.array_equal(out, H)
    for s, d, H in zip(gsynth["lm_h_src"], gsynth["lm_h_dst"], gsynth["lm_h_H"]):
        out = np.zeros(9)
        s = np.ascontiguousarray(s); d = np.ascontiguousarray(d)
        assert hm.hm_run_kernel(_p(s), _p(d), len(s), _p(out)) == 1
        hm.hm_lm_refine(_p(s), _p(d), len(s), _p(out), 10)
        assert np.array_equal(out, H)       # DLT + 9-parameter LM, bit-exact with cv2


def test_pieces_match_oracle(hm, oracle):
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(4, 140))
        s = rng.uniform(0, 640, size=(n, 2)).astype(np.float32)
        d = (s + rng.normal(0, 15, size=(n, 2))).astype(np.float32)
        if rng.random() < 0.3:                       # duplicates -> ill-conditioned
            k = rng.integers(0, n, size=n // 2); s[k] = s[0]; d[k] = d[0]
        H1 = np.zeros(9)
        ok = hm.hm_run_kernel(_p(s), _p(d), n, _p(H1))
        H2 = oracle.run_kernel(s, d)
        assert bool(ok) == (H2 is not None)
        if ok:
            assert np.array_equal(H1, H2.ravel(), equal_nan=True)
            H3 = np.zeros(9)
            assert hm.hm_run_kernel_strided(_p(s), _p(d), n, _p(H3)) == 1
            assert np.array_equal(H1, H3, equal_nan=True)
            if np.isfinite(H1).all():
                Hl = H1.copy(); hm.hm_lm_refine(_p(s), _p(d), n, _p(Hl), 10)
                Ho, _ = oracle.lm_refine(s, d, H2)
                assert np.array_equal(Hl, Ho.ravel(), equal_nan=True)
                err = oracle.compute_error(s, d, H2)
                mine = np.array([hm.hm_reproj_err(_p(H1), *map(float, (s[i, 0], s[i, 1], d[i, 0], d[i, 1]))) for i in range(n)], np.float32)
                assert np.array_equal(err, mine, equal_nan=True)
        a = s[:4].copy(); b = d[:4].copy()
        if rng.random() < 0.3:
            a[3] = a[1]
        assert bool(hm.hm_check_subset4(_p(a), _p(b))) == oracle.check_subset(a, b)


def test_update_num_iters(hm, oracle):
    for n in (5, 17, 32, 100, 8192):
        for good in range(4, n + 1, max(1, n // 97)):
            ep = (n - good) / n
            assert hm.hm_update_num_iters(0.995, ep, 2000) == oracle.update_num_iters(0.995, ep, 4, 2000)


def test_thread_per_hypothesis_run_kernel_bit_exact(hm):
    """csrc/jacobi_thread.cuh (packed indR/indC, predicated static loops, tournament pivot search, strict-upper
    storage) == the serial runKernel of homography_math.cuh, bit for bit, on 4-point samples incl. degenerate ones."""
    rng = np.random.default_rng(11)
    n_ok = 0
    for case in range(4000):
        s = rng.uniform(0, 640, size=(4, 2)).astype(np.float32)
        kind = case % 8
        if kind == 0:
            d = rng.uniform(0, 640, size=(4, 2)).astype(np.float32)            # unrelated points
        elif kind == 1:
            d = s.copy()                                                       # identity
        elif kind == 2:
            d = (s * np.float32(1.3) + np.float32(7)).astype(np.float32)       # similarity
        elif kind == 3:
            s = np.round(s); d = np.round(s + rng.normal(0, 20, size=(4, 2))).astype(np.float32)   # integer coordinates
        elif kind == 4:
            d = (s + rng.normal(0, 15, size=(4, 2))).astype(np.float32); d[1] = d[0]; s[1] = s[0]  # duplicate point
        elif kind == 5:
            s[:, 1] = s[0, 1]; d = (s + rng.normal(0, 5, size=(4, 2))).astype(np.float32)          # zero spread in y
        else:
            d = (s + rng.normal(0, 15, size=(4, 2))).astype(np.float32)
        H1 = np.zeros(9); H2 = np.zeros(9)
        ok1 = hm.hm_run_kernel(_p(s), _p(d), 4, _p(H1))
        ok2 = hm.hm_run_kernel4_thread(_p(s), _p(d), _p(H2))
        assert ok1 == ok2, case
        if ok1:
            assert np.array_equal(H1, H2, equal_nan=True), (case, H1, H2)
            n_ok += 1
    assert n_ok > 3000


def test_foreign_filter_bound_is_the_lattice_predecessor():
    """match_tc_kernel (exact instance) lowers a foreign column part's bound with `foreign - 1.0f` so that equal values
    still pass its strict `>` tests (csrc/match_tc.cu, update_filter).  The accumulator 2 q.t - ||t||^2 of integer
    descriptors in 0..255 over 128 dimensions is an integer with |acc'| <= 2 * 128 * 255^2 < 2^24: on that lattice the
    fp32 subtraction is exact, keeps every value >= the bound and drops every value below it."""
    hi = 2 * 128 * 255 * 255
    assert hi < 2 ** 24
    rng = np.random.default_rng(7)
    v = np.concatenate([rng.integers(-hi, hi + 1, 200000), [-hi, -hi + 1, -1, 0, 1, hi - 1, hi,
                                                           -(2 ** 23), 2 ** 23, 2 ** 23 + 1, -(2 ** 23) - 1]]).astype(np.float32)
    f = (v - np.float32(1.0)).astype(np.float32)
    assert np.all(f.astype(np.float64) == v.astype(np.float64) - 1.0)          # exact
    assert np.all(v > f) and np.all((v - np.float32(1.0)) <= f)                 # equal passes, the next lower integer does not
    assert np.float32(-np.inf) - np.float32(1.0) == np.float32(-np.inf)         # a part that has seen nothing yet
