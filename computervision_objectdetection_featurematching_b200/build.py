"""Builds libcvgraft.so (CUDA, sm_100a) in-tree with nvcc.  No GPU is needed to build."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.environ.get("CVGRAFT_SO") or os.path.join(HERE, "libcvgraft.so")
SOURCES = ["api.cu", "multi.cu", "prep.cu", "match_exact.cu", "match_tc.cu", "ransac.cu"]
HEADERS = ["common.cuh", "ctx.cuh", "homography_math.cuh", "jacobi_warp.cuh", "jacobi_thread.cuh", os.path.join("..", "..", "include", "cvgraft.h")]
# -fmad=false: the verify stage and the exact match kernel must not contract a*b+c (OpenCV's baseline
# build has no FMA); the few fused operations OpenCV does perform are explicit fma() calls.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--shared", "-cudart", "static"]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


OBJ_DIR = os.environ.get("CVGRAFT_OBJ_DIR") or os.path.join(HERE, "build")
COMPILE_FLAGS = [f for f in NVCC_FLAGS if f != "--shared"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return _stale(SO, deps)


def build(force=False, verbose=False):
    """One object per translation unit (compiled side by side, only the stale ones), then one link."""
    if not force and not needs_build():
        return SO
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get("CVG_NVCC_EXTRA", "").split()            # experiments, e.g. -DHYPT_THREADS_DEF=96 -DHYPT_CTAS_DEF=2
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in HEADERS] + [__file__]
    tag = os.path.join(OBJ_DIR, ".flags")
    flags_now = " ".join(COMPILE_FLAGS + extra)
    if force or not os.path.exists(tag) or open(tag).read() != flags_now:
        for f in os.listdir(OBJ_DIR):
            if f.endswith(".o"):
                os.unlink(os.path.join(OBJ_DIR, f))
        with open(tag, "w") as f:
            f.write(flags_now)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not _stale(obj, [path] + hdrs):
            return obj, ""
        cmd = [nvcc()] + COMPILE_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {src}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [nvcc()] + NVCC_FLAGS + ["-o", SO] + [o for o, _ in results] + ["-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libcvgraft.so")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
