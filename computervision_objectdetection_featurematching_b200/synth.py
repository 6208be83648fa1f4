"""Seeded synthetic workloads for parity tests and bench.py.

Shapes follow BASELINE.json's configs; seeds are 1000*config + index.  numpy only.
"""
import numpy as np


def sift_like(rng, n, dim=128):
    """Integer-valued fp32 descriptors in [0,255] with row norm ~512 (what cv2 SIFT emits)."""
    d = np.abs(rng.standard_normal((n, dim), dtype=np.float32))
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    np.minimum(d, 0.2, out=d)
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    return np.minimum(np.round(512.0 * d), 255.0).astype(np.float32)


def float_desc(rng, n, dim=128):
    """Non-integer stress descriptors: the fp32 summation order of cv::batchDistance matters."""
    return (rng.standard_normal((n, dim)) * 37.0).astype(np.float32)


def planted_pair(rng, nq, nt, p_match=0.5, sigma=12.0, dim=128):
    """Query/train descriptor sets where a fraction of queries has a noisy copy among the train rows.
    Returns q [nq,dim], t [nt,dim], truth [nq] (train row of the planted copy or -1)."""
    q = sift_like(rng, nq, dim)
    t = sift_like(rng, nt, dim)
    k = min(int(p_match * nq), nt)
    rows_q = rng.permutation(nq)[:k]
    rows_t = rng.permutation(nt)[:k]
    noisy = q[rows_q] + np.round(rng.normal(0, sigma, size=(k, dim))).astype(np.float32)
    t[rows_t] = np.clip(noisy, 0, 255)
    truth = np.full(nq, -1, np.int32)
    truth[rows_q] = rows_t
    return q, t, truth


def saturated_pair(rng, nq, nt, dim=128):
    """u8-range rows far apart (queries near 0, train near 250): every distance is above 2048, where neighbouring integer
    d^2 round to the same float and cv::BFMatcher's tie rule (lower train index) decides — the regime in which ordering
    by d^2 and ordering by sqrtf(d^2) differ.  Half of the queries own a planted pair of train rows that are their two
    nearest (the rows dip to 130 on the query's 8 signature coordinates) and whose d^2 differ by exactly 1, the larger
    d^2 at the LOWER train index: OpenCV returns that one first whenever the two float distances collide (~12 %)."""
    q = rng.integers(0, 4, size=(nq, dim)).astype(np.float32)
    t = rng.integers(244, 256, size=(nt, dim)).astype(np.float32)
    q[:, dim - 1] = 5
    slots = rng.permutation(nt // 2)[: nq // 2]
    for i, lo in enumerate(slots):
        hi = nt // 2 + int(lo)
        sig = rng.permutation(dim - 1)[:8]
        q[i, sig] = 20
        row = 250.0 - rng.integers(0, 6, size=dim).astype(np.float32)
        row[sig] = 130
        t[hi] = row; t[hi, dim - 1] = 5                    # last coordinate: difference 0
        t[lo] = row; t[lo, dim - 1] = 6                    # difference 1 -> d^2 larger by exactly 1
    return q, t


def random_homography(rng):
    return np.array([[1 + rng.normal(0, .1), rng.normal(0, .1), rng.normal(0, 30)],
                     [rng.normal(0, .1), 1 + rng.normal(0, .1), rng.normal(0, 30)],
                     [rng.normal(0, 2e-4), rng.normal(0, 2e-4), 1.0]])


def correspondences(rng, n, inlier_ratio=0.3, noise=0.7, dup=0.0):
    """n point pairs: a fraction follows a ground-truth homography (+noise px), the rest is uniform.
    dup>0 overwrites that fraction of pairs with copies of other pairs (many-to-one matches, as the
    real dataset has).  Returns src [n,2] f32, dst [n,2] f32, H_true."""
    src = rng.uniform([0, 0], [640, 480], size=(n, 2)).astype(np.float32)
    H = random_homography(rng)
    p = np.c_[src, np.ones(n)] @ H.T
    dst = p[:, :2] / p[:, 2:3] + rng.normal(0, noise, size=(n, 2))
    out = rng.random(n) > inlier_ratio
    dst[out] = rng.uniform([0, 0], [640, 480], size=(int(out.sum()), 2))
    dst = dst.astype(np.float32)
    if dup > 0:
        k = int(n * dup)
        ii = rng.integers(0, n, size=k)
        jj = rng.integers(0, n, size=k)
        src[ii] = src[jj]
        dst[ii] = dst[jj]
    return src, dst, H


def image_pair(rng, nq, nt, p_match=0.5, inlier_ratio=0.3):
    """One synthetic (model view, scene) pair of BASELINE config C3/C4 shape: descriptors with planted
    matches plus keypoints whose planted matches follow a homography for `inlier_ratio` of them."""
    q, t, truth = planted_pair(rng, nq, nt, p_match)
    qk = rng.uniform([0, 0], [640, 480], size=(nq, 2)).astype(np.float32)
    tk = rng.uniform([0, 0], [640, 480], size=(nt, 2)).astype(np.float32)
    H = random_homography(rng)
    m = np.nonzero(truth >= 0)[0]
    geo = m[rng.random(len(m)) < inlier_ratio]
    p = np.c_[qk[geo], np.ones(len(geo))] @ H.T
    tk[truth[geo]] = (p[:, :2] / p[:, 2:3] + rng.normal(0, 0.7, size=(len(geo), 2))).astype(np.float32)
    return q, qk, t, tk, truth, H
