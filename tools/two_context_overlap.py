"""Experiment: two contexts on one GPU driven by two host threads (synchronous calls each) — does the latency-bound
finish kernel of one batch overlap the match / hypothesis kernels of the other?"""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from computervision_objectdetection_featurematching_b200 import api

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NCTX = int(sys.argv[2]) if len(sys.argv) > 2 else 2
STEPS = 24
q, qk, batches = bench.make_workload(3000, B, 4)
ctxs = [api.Context(0) for _ in range(NCTX)]
models = [c.upload_models(q, qk, [0, bench.NQ], [0]) for c in ctxs]
res = [[c.upload_scenes(*b) for b in batches] for c in ctxs]
params = api.detect_params()
def worker(i, n):
    for k in range(n):
        ctxs[i].detect_scenes(models[i], res[i][k % 4], params=params)
for i in range(NCTX):
    worker(i, 3)
torch.cuda.synchronize()
for mode in ("serial", "threads"):
    t0 = time.perf_counter()
    if mode == "serial":
        worker(0, STEPS)
    else:
        th = [threading.Thread(target=worker, args=(i, STEPS // NCTX)) for i in range(NCTX)]
        [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{mode:8s} {NCTX if mode != 'serial' else 1} context(s): {dt / STEPS * 1e3:.3f} ms per step of {B} pairs  -> {B * STEPS / dt:.0f} pairs/s")
