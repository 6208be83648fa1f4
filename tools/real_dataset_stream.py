"""Per-image timing of the streaming real-dataset loop (what bench.py's real_dataset key measures), several passes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data_cache", "features_full.npz"))
ctx = api.Context(0)
md = Z["model_desc"].astype(np.float32); so = Z["scene_offsets"]; n_img = (len(so) - 1) // 5
sd = torch.from_numpy(Z["scene_desc"].astype(np.float32)).pin_memory(); sk = torch.from_numpy(Z["scene_kpt"].astype(np.float32)).pin_memory()
sdn, skn = sd.numpy(), sk.numpy()
models = ctx.upload_models(md, Z["model_kpt"], Z["view_offsets"], Z["view_model"])
sc5 = Z["scales"].astype(np.float32); params = api.detect_params()
def up(i):
    a, b = so[5 * i], so[5 * i + 5]
    return ctx.upload_scenes_async(sdn[a:b], skn[a:b], so[5 * i:5 * i + 6] - a)
for p in range(4):
    t0 = time.perf_counter(); per = []; nxt = up(0)
    for i in range(n_img):
        t1 = time.perf_counter()
        cur = nxt; nxt = up(i + 1) if i + 1 < n_img else None
        t2 = time.perf_counter()
        res, _, _ = ctx.detect_scenes_inliers(models, cur, scales=sc5, params=params)
        t3 = time.perf_counter()
        cur.free()
        per.append((1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (time.perf_counter() - t3)))
    tot = time.perf_counter() - t0
    a = np.array(per)
    print(f"pass {p}: {tot*1e3:.1f} ms  upload-call mean {a[:,0].mean():.2f} max {a[:,0].max():.2f} | detect mean {a[:,1].mean():.2f} max {a[:,1].max():.2f} | free mean {a[:,2].mean():.3f} max {a[:,2].max():.2f}")
