"""BASELINE config 5, RANSAC part only: 2^20 hypotheses on 8192 correspondences at 30 % inliers, no early stop."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, synth
n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
src, dst, _ = synth.correspondences(np.random.default_rng(5002), 8192, 0.3)
ctx = api.Context(0)
ctx.set_timing(True)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter()
    H, mask = ctx.find_homography(src, dst, max_iters=n_hyp, flags=api.RANSAC_NO_EARLY_STOP | api.RANSAC_NO_REFINE)
    dt = time.perf_counter() - t0
    print(f"{n_hyp} hypotheses: {dt*1e3:.1f} ms  {n_hyp/dt/1e6:.2f} M hyps/s  best {int(mask.sum())}  {ctx.last_timing()}")
