"""Where does an e2e step spend its time?  Host-side wall clock of the upload / submit / wait calls of bench.py's e2e loop,
plus variants that isolate the H2D pipeline.  usage: python tools/e2e_probe.py [steps]"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from computervision_objectdetection_featurematching_b200 import api

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
B, R, DEPTH = 64, 4, 3
q, qk, batches = bench.make_workload(3000, B, R)
ctx = api.Context(0)
models = ctx.upload_models(q, qk, [0, bench.NQ], [0])
params = api.detect_params()
pinned = []
for d, k, o in batches:
    pd = torch.from_numpy(d).pin_memory(); pk = torch.from_numpy(k).pin_memory()
    p8 = torch.from_numpy(d.astype(np.uint8)).pin_memory()
    pinned.append((pd.numpy(), pk.numpy(), o, p8.numpy(), pd, pk, p8))


def run(name, upload, detect=True, inliers=True):
    for rep in range(2):
        tu = ts = tw = 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fl = []
        for k in range(steps):
            a = time.perf_counter()
            sc = upload(k % R)
            b = time.perf_counter()
            j = ctx.submit_scenes(models, sc, params=params, want_inliers=inliers) if detect else None
            c = time.perf_counter()
            fl.append((sc, j))
            if len(fl) > DEPTH - 1:
                sc0, j0 = fl.pop(0)
                if j0: j0.wait()
                else: sc0.wait()
                sc0.free()
            d = time.perf_counter()
            tu += b - a; ts += c - b; tw += d - c
        for sc0, j0 in fl:
            if j0: j0.wait()
            else: sc0.wait()
            sc0.free()
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
    print(f"{name:44s} {1e3 * tot / steps:7.3f} ms/step   upload {1e3 * tu / steps:6.3f}  submit {1e3 * ts / steps:6.3f}  wait {1e3 * tw / steps:6.3f}", flush=True)


f32 = lambda j: ctx.upload_scenes_async(pinned[j][0], pinned[j][1], pinned[j][2])
u8 = lambda j: ctx.upload_scenes_u8_async(pinned[j][3], pinned[j][1], pinned[j][2])
if os.environ.get("CVG_TRACE"):
    steps = 8
    run("f32 upload + detect (no inliers)", f32, inliers=False)
    ctx.lib.cvg_trace_dump()
    sys.exit(0)
run("f32 upload only", f32, detect=False)
run("u8 upload only", u8, detect=False)
run("f32 upload + detect (no inliers)", f32, inliers=False)
run("f32 upload + detect + inliers", f32)
run("u8 upload + detect (no inliers)", u8, inliers=False)
run("u8 upload + detect + inliers", u8)
res = [ctx.upload_scenes(d, k, o) for d, k, o in batches]
run("resident + detect (no inliers)", lambda j: res[j], inliers=False) if False else None
ctx.close()
