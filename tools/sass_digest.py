"""SASS / ptxas digest of the tcgen05 match kernels for profiles/: per kernel the ptxas resource line (registers, spills, shared
memory) and the instruction histogram of the cubin — the mnemonics that prove tcgen05 / TMEM / TMA are on the path are listed
first (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit).
    python tools/sass_digest.py > profiles/r02_match_tc_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "computervision_objectdetection_featurematching_b200", "csrc")
obj = "/tmp/match_tc_digest.o"
r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-lineinfo", "-Xcompiler", "-fPIC",
                    "-Xptxas", "-v", "-c", "-o", obj, os.path.join(CSRC, "match_tc.cu")], capture_output=True, text=True, check=True)
res = {}
cur = None
for line in r.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    elif cur and ("registers" in line or "spill" in line):
        res.setdefault(cur, []).append(line.split("ptxas info    :")[-1].strip())
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
hist = {}
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); hist[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        hist[name][m.group(1).split(".")[0]] += 1
key = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "FMNMX3", "FMNMX", "FSETP", "BRA", "BSSY", "BSYNC", "SEL", "FSEL"]
for fn in sorted(hist):
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(f"== {dem}")
    for l in res.get(fn, []):
        print("   ptxas:", l)
    h = hist[fn]
    print(f"   SASS instructions: {sum(h.values())}")
    print("   key mnemonics:", ", ".join(f"{k} x{h[k]}" for k in key if h.get(k)))
    print("   all:", ", ".join(f"{k} x{v}" for k, v in h.most_common()))
    print()
