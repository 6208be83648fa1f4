"""tests/golden/replay_output/ from cv2's own results: the inlier scene points of the 1 308 accepted pairs of the reference's
loop nest, as cv2 4.13.0 returns them (tests/golden/full_dataset_cv2.npz, written by tools/full_dataset_replay.py --cv2),
pushed through the C++ consumer (host/libcvghost_consumer.so: cluster -> box -> merge -> area filter, reference
src/TestsDetector.cpp:112-248) and written in the reference's results layout (src/utils.cpp:12-20, src/Output.cpp:46).
No GPU and no libcvgraft compute is involved: the files are what the reference's own arithmetic + the restated consumer give.

  python tools/make_replay_golden.py [out_dir]     (default: tests/golden/replay_output)
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def consumer():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "libcvghost_consumer.so"])
    return C.CDLL(os.path.join(ROOT, "host", "libcvghost_consumer.so"))


def results_from_points(G, lib):
    """-> {(folder, scene): text of its results file}.  Point order per model = the reference's append order: scales in
    order (:99-109), views of the model in order (:58), inliers in match order (:87-94)."""
    off = G["inlier_offsets"]; xy = G["inlier_xy"]; vm = G["view_model"]
    V = len(vm); n_img = len(G["scene_names"]); n_scales = (len(off) - 1) // V // n_img
    out = {}
    for i in range(n_img):
        lines = []
        for m, name in enumerate(G["model_names"]):
            pts = []
            for k in range(n_scales):
                for v in np.nonzero(vm == m)[0]:
                    pair = (i * n_scales + k) * V + int(v)
                    pts.append(xy[off[pair]:off[pair + 1]])
            p = np.ascontiguousarray(np.concatenate(pts) if pts else np.zeros((0, 2), np.float32), np.float32)
            boxes = np.zeros((64, 4), np.int32)
            n = lib.cvh_consume_points(p.ctypes.data_as(C.c_void_p), len(p), boxes.ctypes.data_as(C.c_void_p), 64)
            for b in boxes[:n]:
                lines.append(f"{name} {b[0]} {b[1]} {b[0] + b[2]} {b[1] + b[3]}\n")
        out[(str(G["model_names"][G["scene_folder"][i]]), str(G["scene_names"][i]))] = "".join(lines)
    return out


if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "replay_output")
    G = np.load(os.path.join(ROOT, "tests", "golden", "full_dataset_cv2.npz"))
    for (folder, scene), text in results_from_points(G, consumer()).items():
        os.makedirs(os.path.join(out_dir, folder), exist_ok=True)
        with open(os.path.join(out_dir, folder, scene + "_results.txt"), "w") as f:
            f.write(text)
    print("wrote", out_dir)
