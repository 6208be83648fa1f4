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
if int(os.environ.get("CVG_TC_EXP", "0")) & 32:
    import ctypes as C
    w = (C.c_int * 64)()
    ctx.lib.cvg_debug_words(ctx.handle, w, 64)
    tw = 64 * 64 * 32 * 16                                 # tile-warps of one launch: pairs x row blocks x tiles x epilogue warps
    print(f"slow-path entries of the previous launch: {w[42]} warp-level = {w[42] / tw:.2f} per tile-warp, {w[43] / max(w[42], 1):.2f} lanes each")
    n = max(w[36], 1)
    print(f"MMA issuer: tiles {w[36]}  cycles per tile waiting for: A tile {64 * w[37] / n:.0f}  accumulator (t_empty) {64 * w[38] / n:.0f}  "
          f"B stages {64 * w[39] / n:.0f}  augmentation {64 * w[40] / n:.0f}   whole loop {64 * w[41] / n:.0f} per tile")
    for name, o in (("warp 0 (part 0)", 20), ("warp 15 (part 3)", 28)):
        n = max(w[o], 1)
        print(f"{name}: tiles {w[o]}  cycles per tile: wait t_full {64 * w[o + 1] / n:.0f}  filter sharing {64 * w[o + 2] / n:.0f}  "
              f"tcgen05.ld + wait {64 * w[o + 3] / n:.0f}  scan + publish {64 * w[o + 4] / n:.0f}   |  per CTA: unit flushes {64 * w[o + 5]} "
              f"unit heads {64 * w[o + 6]}  whole loop {64 * w[o + 7]} cycles")
