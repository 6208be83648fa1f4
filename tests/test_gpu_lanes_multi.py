"""Round-2 structure of the C ABI on the GPU: the d >= 2048 guard of the tensor match path, lanes (sub-batches of one
call, submit / wait), and multi-device contexts (cvg_create_multi: pair sharding + train-tile sharded match).  With one
GPU visible the multi-device tests list that GPU twice (logical devices; the exchange then runs as device-to-device
copies); with >= 2 GPUs they also run over distinct GPUs with the NCCL all-gather."""
import os

import numpy as np
import pytest

from computervision_objectdetection_featurematching_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def api():
    from computervision_objectdetection_featurematching_b200 import api as _api
    return _api


def _n_gpus():
    import torch
    return torch.cuda.device_count()


# ---------------------------------------------------------------- parity hole 4a: sqrtf collisions above d = 2048
def test_match_guard_saturated_rows_vs_cv2(api, oracle):
    Z = np.load(os.path.join(GOLDEN, "golden_sat.npz"))
    q, t = Z["q"].astype(np.float32), Z["t"].astype(np.float32)
    with api.Context(0) as ctx:
        idx, dist, acc = ctx.match_knn2(q, t)
        assert ctx.last_match_path == api.PATH_TENSOR           # integer rows: the tcgen05 kernel serves the call ...
        assert ctx.last_match_guard_rows == len(q)              # ... and every row (all distances >= 2048) is redone exactly
        assert np.array_equal(idx, Z["idx"]) and np.array_equal(dist, Z["dist"]) and np.array_equal(acc, Z["accept"])
        # mixed: mid-range query rows (values 120..139, d ~ 1300 to every train row) next to the saturated ones; only the
        # far rows take the guard
        rng = np.random.default_rng(11)
        qm = rng.integers(120, 140, size=(200, 128)).astype(np.float32)
        q2 = np.concatenate([qm, q[:100]]); t2 = t[:1500]
        got = ctx.match_knn2(q2, t2)
        assert ctx.last_match_path == api.PATH_TENSOR and ctx.last_match_guard_rows == 100
        oi, od = oracle.knn2(q2, t2, nthreads=8)
        assert od[:200].max() < 2048 and od[200:].min() > 2048
        assert np.array_equal(got[0], oi) and np.array_equal(got[1], od) and np.array_equal(got[2], oracle.ratio(oi, od))
        # resident model set + fused call: same guard inside detect (keypoints are irrelevant here)
        models = ctx.upload_models(q2, np.zeros((len(q2), 2), np.float32), [0, 150, len(q2)], [0, 0])
        i3, d3, a3 = ctx.match_knn2(models, t2)
        assert np.array_equal(i3, oi) and np.array_equal(d3, od)
        res, _, _ = ctx.detect_pairs(models, t2, np.zeros((len(t2), 2), np.float32))
        assert res["n_good"].tolist() == [int(oracle.ratio(oi, od)[:150].sum()), int(oracle.ratio(oi, od)[150:].sum())]
        models.free()


# ---------------------------------------------------------------- parity hole 4d: RANSACUpdateNumIters on the host's libm
def test_niters_from_host_libm_equals_device_math(api, gsynth, feats, gpairs):
    """The adaptive iteration count depends on log() / pow(); the device normally uses CUDA's and asks the host only when a
    rounding or cap decision is within 1e-10 of flipping.  With CVG_NITERS_ALL_ON_HOST every evaluation is answered by the
    host's libm (what OpenCV calls): masks, H, iteration counts and gate decisions must not move, on the cv2 goldens."""
    off = gsynth["fh_offsets"]
    with api.Context(0) as c0, api.Context(0, api.NITERS_ALL_ON_HOST) as c1:
        a = c0.find_homography_batch(gsynth["fh_src"], gsynth["fh_dst"], off, want_ransac_mask=True)
        b = c1.find_homography_batch(gsynth["fh_src"], gsynth["fh_dst"], off, want_ransac_mask=True)
        for k in ("H", "mask", "found", "iters", "ransac_mask"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(b["mask"], gsynth["fh_mask"]) and (b["iters"] < 2000).any() and (b["iters"] > 0).any()
        # fused path on the real pairs of the small feature cache (statuses pinned to cv2 by golden_pairs.npz)
        md = feats["model_desc"].astype(np.float32); so = feats["scene_offsets"]
        scales = np.tile(feats["scales"], (len(so) - 1) // 5)
        outs = []
        for c in (c0, c1):
            m = c.upload_models(md, feats["model_kpt"], feats["view_offsets"], feats["view_model"])
            sc = c.upload_scenes(feats["scene_desc"].astype(np.float32), feats["scene_kpt"], so)
            outs.append(c.detect_scenes_inliers(m, sc, scales=scales))
            sc.free(); m.free()
        _same(outs[0], outs[1])
        assert np.array_equal(outs[1][0]["status"], gpairs["status"].astype(np.int32))


# ---------------------------------------------------------------- lanes
def _scene_batch(rng, n_scenes, nq=1500, nt_lo=900, nt_hi=2600):
    q = synth.sift_like(rng, nq)
    qk = rng.uniform([0, 0], [640, 480], size=(nq, 2)).astype(np.float32)
    descs, kpts, off = [], [], [0]
    for _ in range(n_scenes):
        nt = int(rng.integers(nt_lo, nt_hi))
        _, _, t, tk, _, _ = synth.image_pair(rng, nq, nt, p_match=0.5, inlier_ratio=0.5)
        # plant matches of THIS query set
        k = min(nq // 2, nt // 2)
        rq = rng.permutation(nq)[:k]; rt = rng.permutation(nt)[:k]
        t[rt] = np.clip(q[rq] + np.round(rng.normal(0, 10.0, size=(k, 128))).astype(np.float32), 0, 255)
        H = synth.random_homography(rng)
        geo = rng.random(k) < 0.5
        p = np.c_[qk[rq[geo]], np.ones(int(geo.sum()))] @ H.T
        tk[rt[geo]] = (p[:, :2] / p[:, 2:3] + rng.normal(0, 0.7, size=(int(geo.sum()), 2))).astype(np.float32)
        descs.append(t); kpts.append(tk); off.append(off[-1] + nt)
    return q, qk, np.concatenate(descs), np.concatenate(kpts), np.asarray(off, np.int64)


def _same(a, b):
    ra, xa, oa = a; rb, xb, ob = b
    assert ra.tobytes() == rb.tobytes()                          # statuses, counts, iterations, H, det: bit for bit
    assert np.array_equal(oa, ob) and np.array_equal(xa, xb)


def test_lanes_split_and_submit_equal_serial(api):
    """One synchronous call split into sub-batches on 2..4 lanes, and the same batch through submit / wait, give the
    serial call's bytes: per-pair results, inlier points and offsets (sub-batches are views of the resident batch)."""
    rng = np.random.default_rng(2201)
    q, qk, sd, sk, off = _scene_batch(rng, 9)
    vo = [0, 400, 900, 1500]
    scales = np.linspace(0.5, 1.5, 9).astype(np.float32); scales[3] = 1.0
    with api.Context(0) as ctx:
        models = ctx.upload_models(q, qk, vo, [0, 0, 1])
        scenes = ctx.upload_scenes(sd, sk, off)
        ctx.set_lanes(1)
        ref = ctx.detect_scenes_inliers(models, scenes, scales=scales)
        assert (ref[0]["status"] == 0).sum() >= 9               # the planted geometry is found
        os.environ["CVG_SPLIT_MIN_COST"] = "1"                  # split however small the call is
        try:
            for lanes in (2, 3, 4):
                ctx.set_lanes(lanes)
                _same(ctx.detect_scenes_inliers(models, scenes, scales=scales), ref)
                res = ctx.detect_scenes(models, scenes, scales=scales)
                assert res.tobytes() == ref[0].tobytes()
            # pipelined jobs on the lanes: three in flight, waited in order
            jobs = [ctx.submit_scenes(models, scenes, scales=scales) for _ in range(3)]
            for j in jobs:
                _same(j.wait(), ref)
            j = ctx.submit_scenes(models, scenes, scales=scales, want_inliers=False)
            assert j.wait()[0].tobytes() == ref[0].tobytes()
        finally:
            del os.environ["CVG_SPLIT_MIN_COST"]
        # streaming upload + submit (the loop a single-threaded caller runs: src/Output.cpp:27-47)
        ctx.set_lanes(0)
        pending = []
        for k in range(4):
            sc = ctx.upload_scenes_async(sd, sk, off)
            pending.append((sc, ctx.submit_scenes(models, sc, scales=scales)))
            if len(pending) == 3:
                sc0, j0 = pending.pop(0)
                _same(j0.wait(), ref); sc0.free()
        for sc0, j0 in pending:
            _same(j0.wait(), ref); sc0.free()
        scenes.free(); models.free()


def test_lanes_error_is_reported_and_context_survives(api):
    rng = np.random.default_rng(5)
    q, qk, sd, sk, off = _scene_batch(rng, 2, nq=300, nt_lo=300, nt_hi=400)
    with api.Context(0) as ctx:
        models = ctx.upload_models(q, qk, [0, 300], [0])
        scenes = ctx.upload_scenes(sd, sk, off)
        bad = api.detect_params(); bad.ransac.max_iters = 0
        with pytest.raises(api.CvgError):
            ctx.submit_scenes(models, scenes, params=bad)
        res = ctx.submit_scenes(models, scenes).wait()[0]
        assert res.shape == (2, 1)
        scenes.free(); models.free()


# ---------------------------------------------------------------- multi-device context
def _device_lists():
    n = _n_gpus()
    lists = [[0, 0], [0, 0, 0]]                                  # logical devices on one GPU
    if n >= 2:
        lists.append([0, 1])
    if n >= 4:
        lists.append([0, 1, 2, 3])
    return lists


def test_multi_pair_sharding_equals_single_device(api):
    """cvg_create_multi: scenes dealt to the devices by cost, results back in the caller's order: identical bytes for 1,
    2 and 3 (logical) devices — shard-count invariance of the pair-sharded path."""
    rng = np.random.default_rng(2301)
    q, qk, sd, sk, off = _scene_batch(rng, 8)
    vo = [0, 700, 1500]
    scales = np.linspace(0.6, 1.4, 8).astype(np.float32)
    with api.Context(0) as c1:
        c1.set_lanes(1)
        m1 = c1.upload_models(q, qk, vo, [0, 1]); s1 = c1.upload_scenes(sd, sk, off)
        ref = c1.detect_scenes_inliers(m1, s1, scales=scales)
        s1.free(); m1.free()
    for devs in _device_lists():
        with api.Context(devs) as cm:
            assert cm.n_devices == len(devs)
            mm = cm.upload_models(q, qk, vo, [0, 1])
            for up in (cm.upload_scenes, cm.upload_scenes_async):
                sm = up(sd, sk, off)
                _same(cm.detect_scenes_inliers(mm, sm, scales=scales), ref)
                _same(cm.submit_scenes(mm, sm, scales=scales).wait(), ref)
                sm.free()
            sm8 = cm.upload_scenes_u8_async(sd.astype(np.uint8), sk, off)
            _same(cm.detect_scenes_inliers(mm, sm8, scales=scales), ref)
            sm8.free()
            # a small batch (fewer than 2 x n_devices scenes) stays whole on one device; batches take the devices in turn
            o3 = off[:4]
            with api.Context(0) as c1:
                c1.set_lanes(1)
                m1 = c1.upload_models(q, qk, vo, [0, 1]); s1 = c1.upload_scenes(sd[:o3[-1]], sk[:o3[-1]], o3)
                ref3 = c1.detect_scenes_inliers(m1, s1, scales=scales[:3])
                s1.free(); m1.free()
            jobs = []
            for _ in range(len(devs) + 1):
                sc = cm.upload_scenes_async(sd[:o3[-1]], sk[:o3[-1]], o3)
                jobs.append((sc, cm.submit_scenes(mm, sc, scales=scales[:3])))
            for sc, j in jobs:
                _same(j.wait(), ref3); sc.free()
            # the entry points that are not sharded run on the first device
            idx, dist, acc = cm.match_knn2(mm, sd[:off[1]])
            with api.Context(0) as c1:
                i1, d1, a1 = c1.match_knn2(q, sd[:off[1]])
            assert np.array_equal(idx, i1) and np.array_equal(dist, d1) and np.array_equal(acc, a1)
            mm.free()


def test_multi_train_tile_sharded_match(api, oracle):
    """cvg_match_knn2_sharded: per-device local top-2 with global indices, ONE exchange, lexicographic merge — equal to
    the CPU oracle for any number of shards, including ties that straddle shard boundaries and float descriptors."""
    rng = np.random.default_rng(2401)
    q, t, _ = synth.planted_pair(rng, 700, 5000)
    t[2600] = t[100]; q[9] = t[100]; t[4999] = t[3]; q[11] = t[3]      # exact ties across shards
    qf = synth.float_desc(rng, 300, 128); tf = synth.float_desc(rng, 3000, 128)
    tf[2100] = tf[40]; qf[5] = tf[40]
    for devs in _device_lists():
        with api.Context(devs) as cm:
            want = "nccl" if len(set(devs)) == len(devs) else "memcpy"
            assert cm.exchange_kind == want, (devs, cm.exchange_kind)
            for a, b in ((q, t), (qf, tf), (q[:5], t[:300]), (q[:130], t[:2])):
                oi, od = oracle.knn2(a, b, nthreads=8)
                idx, dist, acc = cm.match_knn2_sharded(a, b)
                assert np.array_equal(idx, oi), (devs, a.shape, b.shape)
                assert np.array_equal(dist[oi >= 0], od[oi >= 0])
                assert np.array_equal(acc, oracle.ratio(oi, od))


def test_second_device_context_gets_its_kernel_attributes(api):
    """The dynamic shared memory opt-in is a per-device attribute: a context on GPU 1 of the same process must be able to
    launch the 227 KB kernels too (round-1 advisor finding)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(77)
    q, qk, t, tk, _, _ = synth.image_pair(rng, 600, 1500, p_match=0.5, inlier_ratio=0.5)
    outs = []
    for dev in (0, 1):
        with api.Context(dev) as c:
            m = c.upload_models(q, qk, [0, 600], [0])
            res, inl, off = c.detect_pairs(m, t, tk)
            outs.append((res.tobytes(), inl.tobytes()))
            m.free()
    assert outs[0] == outs[1]
