"""The CPU oracle (oracle/cvoracle.c) against outputs of cv2 4.13.0 itself (tests/golden/, generated
by make_golden.py).  This is the pin that lets the GPU tests trust the oracle."""
import os

import numpy as np
import pytest


def test_eigen_solve_invert_bit_exact(oracle, gsynth):
    for k in range(len(gsynth["eig_n"])):
        n = int(gsynth["eig_n"][k])
        A = gsynth["eig_A"][k][:n, :n]
        W, V = oracle.jacobi(A)
        assert np.array_equal(W, gsynth["eig_w"][k][:n])
        assert np.array_equal(V, gsynth["eig_v"][k][:n, :n])
        assert np.array_equal(oracle.solve_eig(A, gsynth["eig_b"][k][:n]), gsynth["eig_x"][k][:n])
        assert np.array_equal(oracle.invert_eig(A), gsynth["eig_inv"][k][:n, :n])


def test_dlt_bit_exact(oracle, gsynth):
    for s, d, H in zip(gsynth["dlt_src"], gsynth["dlt_dst"], gsynth["dlt_H"]):
        assert np.array_equal(oracle.run_kernel(s, d).ravel(), H)


def test_lm_solver_iterates_bit_exact(oracle, gsynth):
    """cv2.estimateAffine2D(refineIters=k) exposes LMSolver iterate by iterate."""
    for s, d, Hs in zip(gsynth["lm_aff_src"], gsynth["lm_aff_dst"], gsynth["lm_aff_H"]):
        for k in range(1, Hs.shape[0]):
            Hk, _ = oracle.lm_refine_affine(s, d, Hs[0], max_iters=k)
            assert np.array_equal(Hk.ravel(), Hs[k]), k


def test_dlt_plus_lm_bit_exact(oracle, gsynth):
    """findHomography(method=0) = DLT + 9-parameter LM refine."""
    for s, d, H in zip(gsynth["lm_h_src"], gsynth["lm_h_dst"], gsynth["lm_h_H"]):
        r = oracle.find_homography(s, d, method=0)
        assert np.array_equal(r["H"].ravel(), H)


def test_knn_float_and_ties(oracle, gsynth):
    for tag in ("float", "int"):
        idx, dist = oracle.knn2(gsynth[f"knn_{tag}_q"], gsynth[f"knn_{tag}_t"], nthreads=4)
        assert np.array_equal(idx, gsynth[f"knn_{tag}_idx"])
        assert np.array_equal(dist, gsynth[f"knn_{tag}_dist"])


def test_knn_saturated_rows_sqrt_collisions(oracle):
    """Distances above 2048: distinct integer d^2 collapse into one float and cv2's tie rule picks the lower index
    (golden_sat.npz: 78 of 300 rows differ from the d^2 order).  The oracle orders by the rounded distance like cv2."""
    Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_sat.npz"))
    q, t = Z["q"].astype(np.float32), Z["t"].astype(np.float32)
    idx, dist = oracle.knn2(q, t, nthreads=4)
    assert np.array_equal(idx, Z["idx"]) and np.array_equal(dist, Z["dist"])
    assert np.array_equal(oracle.ratio(idx, dist), Z["accept"])
    assert dist.min() > 2048 and (dist[:, 0] == dist[:, 1]).sum() > 50


def test_knn_tiny_train_sets(oracle, gsynth):
    q, t = gsynth["knn_int_q"], gsynth["knn_int_t"]
    for nt in (1, 2):
        idx, dist = oracle.knn2(q[:9], t[:nt])
        gi, gd = gsynth[f"knn_nt{nt}_idx"], gsynth[f"knn_nt{nt}_dist"]
        assert np.array_equal(idx, gi)
        assert np.array_equal(dist[gi >= 0], gd[gi >= 0])
    idx, _ = oracle.knn2(q[:9], t[:0])
    assert (idx == -1).all()
    assert not oracle.ratio(idx, np.zeros_like(idx, dtype=np.float32)).any()


def _check_fh(oracle, src, dst, found, H, mask, **kw):
    r = oracle.find_homography(src, dst, **kw)
    assert r["found"] == bool(found)
    assert np.array_equal(r["mask"], mask)            # bit-exact inlier mask
    if found:
        rel = np.max(np.abs(r["H"].ravel() - H) / np.maximum(np.abs(H), 1e-12))
        assert rel < 1e-5, rel
    return r


def test_find_homography_synthetic(oracle, gsynth):
    off = gsynth["fh_offsets"]
    exact = 0
    for k in range(len(off) - 1):
        a, b = off[k], off[k + 1]
        r = _check_fh(oracle, gsynth["fh_src"][a:b], gsynth["fh_dst"][a:b], gsynth["fh_found"][k],
                      gsynth["fh_H"][k], gsynth["fh_mask"][a:b])
        exact += bool(gsynth["fh_found"][k]) and np.array_equal(r["H"].ravel(), gsynth["fh_H"][k])
    # below 50 inliers every stage is pinned bit-for-bit; above, OpenCV hands J^T r to OpenBLAS
    assert exact >= 45


def test_find_homography_params(oracle, gsynth):
    for (thr, it, conf), H, mask in zip(gsynth["fhp_params"], gsynth["fhp_H"], gsynth["fhp_mask"]):
        _check_fh(oracle, gsynth["fhp_src"], gsynth["fhp_dst"], True, H, mask,
                  thr=float(thr), max_iters=int(it), conf=float(conf))


def test_find_homography_rejects_short_input(oracle):
    with pytest.raises(ValueError):
        oracle.find_homography(np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32))


def test_real_pairs_match_stage(oracle, feats, gpairs):
    """knnMatch + ratio test on real SIFT features: indices, distances and accept mask bit-exact."""
    md = feats["model_desc"].astype(np.float32)
    so = feats["scene_offsets"]
    for s in range(0, len(so) - 1, 3):
        t = feats["scene_desc"][so[s]:so[s + 1]].astype(np.float32)
        idx, dist = oracle.knn2(md, t, nthreads=os.cpu_count())
        assert np.array_equal(idx, gpairs["idx"][s].astype(np.int32))
        assert np.array_equal(dist, gpairs["dist"][s])
        assert np.array_equal(oracle.ratio(idx, dist), gpairs["accept"][s])


def test_real_pairs_verify_stage(oracle, feats, gpairs):
    """findHomography + gates on every golden real pair (ill-conditioned sets with duplicate
    correspondences): status, inlier mask bit-exact; H within 1e-5 relative."""
    mk = feats["model_kpt"]; vo = feats["view_offsets"]; so = feats["scene_offsets"]
    V = len(vo) - 1
    moff = gpairs["mask_offsets"]
    n_pairs = 0; n_exact = 0
    for s in range(len(so) - 1):
        tk = feats["scene_kpt"][so[s]:so[s + 1]]
        idx = gpairs["idx"][s].astype(np.int32); acc = gpairs["accept"][s]
        for v in range(0, V, 2):
            sel = np.nonzero(acc[vo[v]:vo[v + 1]])[0] + vo[v]
            p = s * V + v
            gmask = gpairs["mask"][moff[p]:moff[p + 1]]
            gst = int(gpairs["status"][s, v])
            assert len(sel) == gpairs["n_good"][s, v]
            if len(sel) < 4:
                assert gst == oracle.LT4_MATCHES
                continue
            src = mk[sel]; dst = tk[idx[sel, 0]]
            r = oracle.find_homography(src, dst)
            st = oracle.gate(len(sel), r["found"], r["mask"], r["H"])
            assert st == gst, (s, v)
            assert np.array_equal(r["mask"], gmask), (s, v)
            if r["found"]:
                H = gpairs["H"][s, v]
                rel = np.max(np.abs(r["H"].ravel() - H) / np.maximum(np.abs(H), 1e-12))
                assert rel < 1e-5, (s, v, rel)
                n_exact += np.array_equal(r["H"].ravel(), H)
            n_pairs += 1
    assert n_pairs > 300 and n_exact >= 0.9 * n_pairs
