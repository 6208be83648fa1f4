"""Small bench-shaped run for ncu: one resident 8192-row view against 8 resident scenes of 8192 rows."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from computervision_objectdetection_featurematching_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
q, qk, batches = bench.make_workload(3000, B, 1)
ctx = api.Context(0); ctx.set_timing(True)
models = ctx.upload_models(q, qk, [0, bench.NQ], [0])
sc = ctx.upload_scenes(*batches[0])
for _ in range(3):
    res = ctx.detect_scenes(models, sc)
    print(ctx.last_timing(), int((res["status"] == 0).sum()))
