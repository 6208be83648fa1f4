import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, synth
rng = np.random.default_rng(1)
ctx = api.Context(0)
for n in (32, 8192):
    s, d, _ = synth.correspondences(rng, n, 0.3 if n > 100 else 0.5)
    for _ in range(2):
        ctx.find_homography(s, d)
print("done")
