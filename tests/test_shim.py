"""shim/cvgraft_opencv.hpp — the reference-side C++ binding (INTEGRATION.md as code).  OpenCV C++ is not installed here, so
the header is compiled against a minimal stand-in for the cv:: types it touches (tests/shim/opencv_stub); the GPU test
runs both forms the way the patched src/TestsDetector.cpp would — call for call (knnMatch + ratio test + findHomography
+ gates) and fused — on real cached features and checks that they append the same inlier points and agree with the
cv2 golden gate decisions."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "computervision_objectdetection_featurematching_b200")
EXE = os.path.join(ROOT, "tests", "shim", "shim_main")


def build_driver():
    from computervision_objectdetection_featurematching_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "shim", "shim_main.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror",
                           "-I" + os.path.join(ROOT, "tests", "shim", "opencv_stub"), "-I" + os.path.join(ROOT, "tests", "shim"),
                           "-I" + os.path.join(ROOT, "shim"), "-I" + os.path.join(ROOT, "include"), "-o", EXE, src,
                           "-L" + PKG, "-lcvgraft", "-Wl,-rpath," + PKG])
    return EXE


def test_shim_compiles_against_the_c_abi():
    assert os.path.exists(build_driver())


@pytest.mark.gpu
def test_shim_call_for_call_equals_fused_and_golden(feats, gpairs, tmp_path):
    exe = EXE if os.path.exists(EXE) else build_driver()
    vo = feats["view_offsets"].astype(np.int32); so = feats["scene_offsets"]
    md = feats["model_desc"].astype(np.float32); mk = feats["model_kpt"].astype(np.float32)
    for s in (2, 6):                                           # two (image, scale) scenes of the small cache
        scale = float(feats["scales"][s % 5])
        sd = feats["scene_desc"][so[s]:so[s + 1]].astype(np.float32); sk = feats["scene_kpt"][so[s]:so[s + 1]].astype(np.float32)
        inp = tmp_path / f"in{s}.bin"; out = tmp_path / f"out{s}.bin"
        with open(inp, "wb") as f:
            f.write(struct.pack("<iii", len(vo) - 1, len(sd), int(round(scale * 1000))))
            f.write(vo.tobytes()); f.write(md.tobytes()); f.write(mk.tobytes()); f.write(sd.tobytes()); f.write(sk.tobytes())
        r = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = open(out, "rb").read()
        n1, n2 = struct.unpack("<ii", raw[:8])
        a = np.frombuffer(raw[8:8 + 8 * n1], np.float32).reshape(-1, 2)
        b = np.frombuffer(raw[8 + 8 * n1:], np.float32).reshape(-1, 2)
        assert n1 == n2 and np.array_equal(a, b)               # the two forms append the same points, bit for bit
        st = gpairs["status"][s]; ni = gpairs["n_inliers"][s]
        assert n1 == int(ni[st == 0].sum())                    # = the inliers of the views cv2 accepts
