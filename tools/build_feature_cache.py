"""Build the feature cache: run the reference's *plumbing* (OpenCV SIFT on masked model views and on
5-scale test scenes; reference src/ModelsDetector.cpp:47-80, src/TestsDetector.cpp:33,99-106) once
with cv2, and store descriptors/keypoints as flat arrays so the hot path can be replayed without
OpenCV.  Needs /root/reference (data) and cv2 — runs in the build container only.

Iteration order: the reference iterates fs::directory_iterator / unordered_map (unspecified order);
the cache uses sorted order (SURVEY.md §6 "sorted iteration order").

Output (np.savez_compressed):
  model_desc  u8 [N,128]   all model-view descriptors (SIFT values are integers 0..255)
  model_kpt   f32 [N,2]
  view_offsets i32 [V+1], view_model i32 [V], model_names, view_names
  scene_desc  u8 [M,128], scene_kpt f32 [M,2], scene_offsets i32 [S*5+1],
  scene_names [S], scene_folder [S], scales f32 [5]
"""
import os
import sys

import cv2
import numpy as np

SCALES = np.array([0.7, 0.85, 1.0, 1.15, 1.3], np.float32)
FOLDERS = ["004_sugar_box", "006_mustard_bottle", "035_power_drill"]


def build(data_root, out_path, scenes_per_folder=None):
    sift = cv2.SIFT_create()
    mdesc, mkpt, voff, vmodel, vnames = [], [], [0], [], []
    for mi, folder in enumerate(FOLDERS):
        mdir = os.path.join(data_root, folder, "models")
        files = sorted(os.listdir(mdir))
        color = {f[:f.find("_color")]: f for f in files if "_color" in f}
        mask = {f[:f.find("_mask")]: f for f in files if "_mask" in f and "_color" not in f}
        for base in sorted(color):
            img = cv2.imread(os.path.join(mdir, color[base]), cv2.IMREAD_GRAYSCALE)
            if img is None:
                continue
            m = cv2.imread(os.path.join(mdir, mask[base]), cv2.IMREAD_GRAYSCALE) if base in mask else None
            kp, des = sift.detectAndCompute(img, m)
            if des is None:
                des = np.zeros((0, 128), np.float32)
            assert np.all(des == np.round(des)) and des.min(initial=0) >= 0 and des.max(initial=0) <= 255
            mdesc.append(des.astype(np.uint8))
            mkpt.append(np.array([k.pt for k in kp], np.float32).reshape(-1, 2))
            voff.append(voff[-1] + len(kp)); vmodel.append(mi); vnames.append(f"{folder}/{base}")
    sdesc, skpt, soff, snames, sfolder = [], [], [0], [], []
    for fi, folder in enumerate(FOLDERS):
        tdir = os.path.join(data_root, folder, "test_images")
        files = sorted(f for f in os.listdir(tdir) if f.endswith((".jpg", ".png")))
        if scenes_per_folder is not None:
            files = files[:scenes_per_folder]
        for f in files:
            scene = cv2.imread(os.path.join(tdir, f), cv2.IMREAD_COLOR)
            gray = cv2.cvtColor(scene, cv2.COLOR_BGR2GRAY)
            for s in SCALES:
                scaled = cv2.resize(gray, None, fx=float(s), fy=float(s))
                kp, des = sift.detectAndCompute(scaled, None)
                if des is None:
                    des = np.zeros((0, 128), np.float32)
                assert np.all(des == np.round(des))
                sdesc.append(des.astype(np.uint8))
                skpt.append(np.array([k.pt for k in kp], np.float32).reshape(-1, 2))
                soff.append(soff[-1] + len(kp))
            snames.append(os.path.splitext(f)[0]); sfolder.append(fi)
    np.savez_compressed(
        out_path,
        model_desc=np.concatenate(mdesc), model_kpt=np.concatenate(mkpt),
        view_offsets=np.array(voff, np.int32), view_model=np.array(vmodel, np.int32),
        model_names=np.array(FOLDERS), view_names=np.array(vnames),
        scene_desc=np.concatenate(sdesc), scene_kpt=np.concatenate(skpt),
        scene_offsets=np.array(soff, np.int32), scene_names=np.array(snames),
        scene_folder=np.array(sfolder, np.int32), scales=SCALES)
    return out_path


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "data_cache", "features_full.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    n = int(sys.argv[3]) if len(sys.argv) > 3 else None
    print(build(root, out, n))
