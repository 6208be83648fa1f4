"""Builds libcvgraft.so (CUDA, sm_100a) in-tree with nvcc.  No GPU is needed to build."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcvgraft.so")
SOURCES = ["api.cu", "prep.cu", "match_exact.cu", "match_tc.cu", "ransac.cu"]
HEADERS = ["common.cuh", "homography_math.cuh", "jacobi_warp.cuh", "jacobi_thread.cuh", os.path.join("..", "..", "include", "cvgraft.h")]
# -fmad=false: the verify stage and the exact match kernel must not contract a*b+c (OpenCV's baseline
# build has no FMA); the few fused operations OpenCV does perform are explicit fma() calls.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--shared", "-cudart", "static"]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    extra = os.environ.get("CVG_NVCC_EXTRA", "").split()            # experiments, e.g. -DHYPT_THREADS_DEF=96 -DHYPT_CTAS_DEF=2
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libcvgraft.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
