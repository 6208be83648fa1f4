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


def run_cv2(Z):
    import cv2
    from multiprocessing import Pool
    md = Z["model_desc"].astype(np.float32); mk = Z["model_kpt"]; vo = Z["view_offsets"]; so = Z["scene_offsets"]
    S = len(so) - 1; V = len(vo) - 1
    out = np.zeros((S, V), dtype=[("status", "<i4"), ("n_good", "<i4"), ("n_inliers", "<i4"), ("H", "<f8", (9,))])
    bf = cv2.BFMatcher(cv2.NORM_L2)
    t0 = time.time()
    for s in range(S):
        t = Z["scene_desc"][so[s]:so[s + 1]].astype(np.float32); tk = Z["scene_kpt"][so[s]:so[s + 1]]
        m = bf.knnMatch(md, t, 2)
        idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m]); dist = np.array([[a.distance, b.distance] for a, b in m], np.float32)
        acc = dist[:, 0] < np.float32(0.9) * dist[:, 1]
        for v in range(V):
            sel = np.nonzero(acc[vo[v]:vo[v + 1]])[0] + vo[v]
            out["n_good"][s, v] = len(sel)
            if len(sel) < 4:
                out["status"][s, v] = 1; continue
            H, mask = cv2.findHomography(mk[sel], tk[idx[sel, 0]], cv2.RANSAC, 5.0)
            if H is None:
                out["status"][s, v] = 2; continue
            out["H"][s, v] = H.ravel(); out["n_inliers"][s, v] = int(mask.sum())
            if mask.sum() < 4:
                out["status"][s, v] = 3; continue
            d = abs(np.linalg.det(H))
            m0 = H.ravel(); d = abs(m0[0] * (m0[4] * m0[8] - m0[5] * m0[7]) - m0[1] * (m0[3] * m0[8] - m0[5] * m0[6]) + m0[2] * (m0[3] * m0[7] - m0[4] * m0[6]))
            out["status"][s, v] = 4 if (d < np.float32(0.1) or d > np.float32(10.0)) else 0
        print(f"scene-scale {s + 1}/{S}  {time.time() - t0:.0f}s", flush=True)
    return out


if __name__ == "__main__":
    Z = load()
    if "--cv2" in sys.argv:
        out = run_cv2(Z)
        np.savez_compressed(os.path.join(ROOT, "data_cache", "replay_cv2.npz"), res=out)
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
