"""Export a feature cache (.npz from build_feature_cache.py / make_golden.py) to the flat binary the C++ host
driver reads (host/feature_cache.hpp).  Layout, little endian:
  char magic[4]="CVGF"; int32 version=1, n_views, n_scenes, n_scales, n_model_rows; int64 n_scene_rows
  int32 view_offsets[V+1]; int32 view_model[V]; char model_names[3][64]
  int64 scene_offsets[S*n_scales+1]; int32 scene_folder[S]; char scene_names[S][64]; float scales[n_scales]
  uint8 model_desc[N][128]; float model_kpt[N][2]; uint8 scene_desc[M][128]; float scene_kpt[M][2]
"""
import struct
import sys

import numpy as np


def export(npz, out):
    Z = np.load(npz)
    V = len(Z["view_offsets"]) - 1; S = len(Z["scene_names"]); K = len(Z["scales"])
    N = Z["model_desc"].shape[0]; M = Z["scene_desc"].shape[0]
    with open(out, "wb") as f:
        f.write(b"CVGF"); f.write(struct.pack("<iiiiiq", 1, V, S, K, N, M))
        f.write(Z["view_offsets"].astype("<i4").tobytes()); f.write(Z["view_model"].astype("<i4").tobytes())
        for n in Z["model_names"]:
            f.write(str(n).encode()[:63].ljust(64, b"\0"))
        f.write(Z["scene_offsets"].astype("<i8").tobytes()); f.write(Z["scene_folder"].astype("<i4").tobytes())
        for n in Z["scene_names"]:
            f.write(str(n).encode()[:63].ljust(64, b"\0"))
        f.write(Z["scales"].astype("<f4").tobytes())
        f.write(np.ascontiguousarray(Z["model_desc"], np.uint8).tobytes()); f.write(Z["model_kpt"].astype("<f4").tobytes())
        f.write(np.ascontiguousarray(Z["scene_desc"], np.uint8).tobytes()); f.write(Z["scene_kpt"].astype("<f4").tobytes())
    return out


if __name__ == "__main__":
    print(export(sys.argv[1], sys.argv[2]))
