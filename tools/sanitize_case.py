"""Smallest case that touches every kernel once (for compute-sanitizer --tool memcheck): integer match, float match with
fallback rows, forced SIMT match, batched findHomography through the thread-per-hypothesis kernel, the fused path with the
streaming upload."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, synth

rng = np.random.default_rng(3)
ctx = api.Context(0)
q, qk, t, tk, truth, H = synth.image_pair(rng, 300, 700, p_match=0.5, inlier_ratio=0.5)
idx, dist, acc = ctx.match_knn2(q, t); print("int path", ctx.last_match_path, int(acc.sum()))
idx_int = idx.copy()
qf = synth.float_desc(rng, 200, 128); tf = synth.float_desc(rng, 900, 128); tf[500:520] = tf[:20] + np.float32(1e-3)
idx, dist, acc = ctx.match_knn2(qf, tf); print("float path", ctx.last_match_path, "fallback", ctx.last_match_fallback_rows)
cx = api.Context(0, api.FORCE_EXACT_MATCH); cx.match_knn2(qf[:130], tf[:300]); print("simt path", cx.last_match_path); cx.close()
srcs, dsts, offs = [], [], [0]
for n in [int(v) for v in rng.integers(4, 120, size=40)] + [3000]:
    s, d, _ = synth.correspondences(rng, n, 0.5)
    srcs.append(s); dsts.append(d); offs.append(offs[-1] + n)
out = ctx.find_homography_batch(np.concatenate(srcs), np.concatenate(dsts), offs, max_iters=400)
print("batch found", int(out["found"].sum()))
Hh, m = ctx.find_homography(srcs[-1], dsts[-1]); print("single", Hh is not None, int(m.sum()))
models = ctx.upload_models(q, qk, [0, 100, 300], [0, 0])
sc = ctx.upload_scenes_async(np.concatenate([t, t[:300]]), np.concatenate([tk, tk[:300]]), [0, 700, 1000])
res = ctx.detect_scenes(models, sc, scales=[1.0, 1.15]); print("detect", res["status"].tolist())
r2, inl, ioff = ctx.detect_pairs(models, t, tk, scale=0.85); print("pairs", r2["status"].tolist(), len(inl))
os.environ["CVG_SPLIT_MIN_COST"] = "1"                   # lanes: sub-batches of one call, then the asynchronous pair
r3 = ctx.detect_scenes(models, sc, scales=[1.0, 1.15]); print("lanes", bool(r3.tobytes() == res.tobytes()))
r4, _, _ = ctx.submit_scenes(models, sc, scales=[1.0, 1.15]).wait(); print("submit", bool(r4.tobytes() == res.tobytes()))
qs, ts = synth.saturated_pair(rng, 60, 400)              # distances above 2048: the guard's exact kernels
ctx.match_knn2(qs, ts); print("guard rows", ctx.last_match_guard_rows)
sc.free(); models.free(); ctx.close()
cm = api.Context([0, 0])                                 # two logical devices: dealing, device-to-device exchange, merge
mm = cm.upload_models(q, qk, [0, 100, 300], [0, 0])
sm = cm.upload_scenes(np.concatenate([t, t[:300], t[:500], t[:256]]), np.concatenate([tk, tk[:300], tk[:500], tk[:256]]), [0, 700, 1000, 1500, 1756])
print("multi", cm.detect_scenes(mm, sm)["status"].tolist())
i5, d5, a5 = cm.match_knn2_sharded(q, t); print("sharded", cm.exchange_kind, bool(np.array_equal(i5, idx_int)))
sm.free(); mm.free(); cm.close()
print("sanitize case done")
