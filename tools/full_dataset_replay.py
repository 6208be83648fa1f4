"""Full-dataset replay of the reference's hot loop nest (30 test images x 5 scales x 89 views = 13 350 pairs,
src/TestsDetector.cpp:38,58,99-100) through libcvgraft, from the feature cache (data_cache/features_full.npz,
built by tools/build_feature_cache.py from the reference's data with cv2 SIFT).

  python tools/full_dataset_replay.py            -> GPU run, prints gate histogram + timing, writes gpurun_out/replay_gpu.npz
  python tools/full_dataset_replay.py --cv2      -> same loop nest with cv2 (the reference's arithmetic), writes
                                                     data_cache/replay_cv2.npz (slow: ~10 min on 8 cores)
  python tools/full_dataset_replay.py --compare  -> compares the two
"""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
CACHE = os.path.join(ROOT, "data_cache", "features_full.npz")


def load():
    Z = np.load(CACHE)
    return {k: Z[k] for k in Z.files}


def run_gpu(Z):
    from computervision_objectdetection_featurematching_b200 import api
    md = Z["model_desc"].astype(np.float32)
    so = Z["scene_offsets"]
    S = len(so) - 1
    scales = np.tile(Z["scales"], S // 5).astype(np.float32)
    with api.Context(0) as ctx:
        ctx.set_timing(True)
        models = ctx.upload_models(md, Z["model_kpt"], Z["view_offsets"], Z["view_model"])
        t0 = time.perf_counter()
        scenes = ctx.upload_scenes(Z["scene_desc"].astype(np.float32), Z["scene_kpt"], so)
        t1 = time.perf_counter()
        res = ctx.detect_scenes(models, scenes, scales=scales)
        t2 = time.perf_counter()
        res = ctx.detect_scenes(models, scenes, scales=scales)
        t3 = time.perf_counter()
        tm = ctx.last_timing()
        print(f"pairs {res.size}  upload {1e3*(t1-t0):.1f} ms  detect(first) {1e3*(t2-t1):.1f} ms  detect(warm) {1e3*(t3-t2):.1f} ms  {tm}")
        print("pairs/s (warm, resident):", res.size / (t3 - t2))
    return res


_Z = None


def _cv2_scene(s):
    """One scaled scene against all 89 views with cv2 — the reference's arithmetic for src/TestsDetector.cpp:58-95, incl.
    the inlier mask (:78) and the inlier scene points divided by the scale (:87-94, :48-55)."""
    import cv2
    cv2.setNumThreads(1)
    Z = _Z
    md = Z["model_desc"].astype(np.float32); mk = Z["model_kpt"]; vo = Z["view_offsets"]; so = Z["scene_offsets"]
    V = len(vo) - 1
    scale = np.float32(Z["scales"][s % 5])
    out = np.zeros(V, dtype=[("status", "<i4"), ("n_good", "<i4"), ("n_inliers", "<i4"), ("H", "<f8", (9,))])
    masks, pts = [], []
    t = Z["scene_desc"][so[s]:so[s + 1]].astype(np.float32); tk = Z["scene_kpt"][so[s]:so[s + 1]].astype(np.float32)
    m = cv2.BFMatcher(cv2.NORM_L2).knnMatch(md, t, 2)
    idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m]); dist = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    acc = dist[:, 0] < np.float32(0.9) * dist[:, 1]
    for v in range(V):
        sel = np.nonzero(acc[vo[v]:vo[v + 1]])[0] + vo[v]
        out["n_good"][v] = len(sel)
        mask = np.zeros(len(sel), np.uint8); inl = np.zeros((0, 2), np.float32)
        if len(sel) < 4:
            out["status"][v] = 1
        else:
            H, mk_ = cv2.findHomography(mk[sel], tk[idx[sel, 0]], cv2.RANSAC, 5.0)
            if H is None:
                out["status"][v] = 2
            else:
                mask = mk_.ravel().astype(np.uint8)
                out["H"][v] = H.ravel(); out["n_inliers"][v] = int(mask.sum())
                if mask.sum() < 4:
                    out["status"][v] = 3
                else:
                    m0 = H.ravel(); d = abs(m0[0] * (m0[4] * m0[8] - m0[5] * m0[7]) - m0[1] * (m0[3] * m0[8] - m0[5] * m0[6]) + m0[2] * (m0[3] * m0[7] - m0[4] * m0[6]))
                    out["status"][v] = 4 if (d < np.float32(0.1) or d > np.float32(10.0)) else 0
                    if out["status"][v] == 0:
                        inl = tk[idx[sel, 0]][mask != 0].astype(np.float32)
                        if scale != np.float32(1.0):
                            inl = (inl / scale).astype(np.float32)          # fp32 divide, as scalePoints does
        masks.append(mask); pts.append(inl)
    return s, out, masks, pts


def run_cv2(Z):
    """-> per-pair records [S, V], the returned inlier masks of every pair (concatenated, pair order) and the inlier scene
    points of the ACCEPTED pairs as the reference appends them (concatenated, pair order)."""
    global _Z
    from multiprocessing import Pool
    _Z = Z
    S = len(Z["scene_offsets"]) - 1; V = len(Z["view_offsets"]) - 1
    out = np.zeros((S, V), dtype=[("status", "<i4"), ("n_good", "<i4"), ("n_inliers", "<i4"), ("H", "<f8", (9,))])
    masks = [None] * S; pts = [None] * S
    t0 = time.time()
    with Pool(os.cpu_count()) as pool:
        for k, (s, o, mk, pt) in enumerate(pool.imap_unordered(_cv2_scene, range(S))):
            out[s] = o; masks[s] = mk; pts[s] = pt
            print(f"scene-scale {k + 1}/{S}  {time.time() - t0:.0f}s", flush=True)
    mask_off = np.zeros(S * V + 1, np.int64); inl_off = np.zeros(S * V + 1, np.int64)
    for s in range(S):
        for v in range(V):
            mask_off[s * V + v + 1] = mask_off[s * V + v] + len(masks[s][v])
            inl_off[s * V + v + 1] = inl_off[s * V + v] + len(pts[s][v])
    mask_cat = np.concatenate([m for ms in masks for m in ms]) if S else np.zeros(0, np.uint8)
    inl_cat = np.concatenate([p for ps in pts for p in ps]).astype(np.float32) if S else np.zeros((0, 2), np.float32)
    return out, mask_cat, mask_off, inl_cat, inl_off


if __name__ == "__main__":
    Z = load()
    if "--cv2" in sys.argv:
        out, mask_cat, mask_off, inl_cat, inl_off = run_cv2(Z)
        np.savez_compressed(os.path.join(ROOT, "data_cache", "replay_cv2.npz"), res=out)
        # the committed golden: gate records + returned masks + the accepted pairs' inlier scene points (src/TestsDetector.cpp:87-94)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "full_dataset_cv2.npz"), status=out["status"].astype(np.int8),
                            n_good=out["n_good"], n_inliers=out["n_inliers"], H=out["H"], mask=np.packbits(mask_cat),
                            mask_offsets=mask_off, inlier_xy=inl_cat, inlier_offsets=inl_off,
                            # what the consumer + results writer need besides the points (tools/make_replay_golden.py)
                            view_model=Z["view_model"], model_names=Z["model_names"], scene_names=Z["scene_names"],
                            scene_folder=Z["scene_folder"])
    elif "--compare" in sys.argv:
        a = np.load(os.path.join(ROOT, "data_cache", "replay_cv2.npz"))["res"]
        b = np.load(os.path.join(ROOT, "gpurun_out", "replay_gpu.npz"))["res"]
        print("pairs", a.size, "status equal:", int((a["status"] == b["status"]).sum()), "n_good equal:", int((a["n_good"] == b["n_good"]).sum()))
        ok = np.isin(a["status"], (0, 3, 4))
        print("n_inliers equal (where H exists):", int((a["n_inliers"][ok] == b["n_inliers"][ok]).sum()), "of", int(ok.sum()))
        rel = np.abs(a["H"][ok] - b["H"][ok]) / np.maximum(np.abs(a["H"][ok]), 1e-12)
        print("H max rel diff: %.3e   bit-exact H: %d of %d" % (rel.max(), int((a["H"][ok] == b["H"][ok]).all(1).sum()), int(ok.sum())))
        print("cv2 gate histogram:", np.bincount(a["status"].ravel(), minlength=5).tolist(), " gpu:", np.bincount(b["status"].ravel(), minlength=5).tolist())
    else:
        res = run_gpu(Z)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "replay_gpu.npz"), res=res)
        print("gate histogram [accept, <4 matches, H empty, <4 inliers, det]:", np.bincount(res["status"].ravel(), minlength=5).tolist())
