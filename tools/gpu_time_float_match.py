"""Wall time of cvg_match_knn2 (8192 resident queries vs one 8192-row train matrix from host memory) for the three
match paths: integer descriptors (tensor, exact), float descriptors (tensor candidates + re-rank), forced SIMT kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, synth

rng = np.random.default_rng(1)
qi, ti, _ = synth.planted_pair(rng, 8192, 8192)
qf = synth.float_desc(rng, 8192, 128); tf = synth.float_desc(rng, 8192, 128)
rows = rng.permutation(8192)[:4096]
tf[rows[:2048]] = qf[rows[2048:]] + rng.normal(0, 6.0, size=(2048, 128)).astype(np.float32)
kp = np.zeros((8192, 2), np.float32)
for name, flags, q, t in (("integer/tensor", 0, qi, ti), ("float/tensor+rerank", 0, qf, tf), ("float/simt", api.FORCE_EXACT_MATCH, qf, tf)):
    ctx = api.Context(0, flags); ctx.set_timing(True)
    m = ctx.upload_models(q, kp, [0, 8192], [0])
    for _ in range(3):
        ctx.match_knn2(m, t)
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.match_knn2(m, t)
    dt = (time.perf_counter() - t0) / 10
    print(f"{name:22s} path={ctx.last_match_path} call {dt*1e3:.3f} ms  match-kernel {ctx.last_timing()['match_ms']*1e3:.1f} us"
          f"  fallback rows {ctx.last_match_fallback_rows}")
    m.free(); ctx.close()
