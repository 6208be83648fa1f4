"""First-contact diagnostic for a GPU box: runs each stage once and prints what differs from the oracle."""
import sys, os, time, traceback
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, synth
from oracle import cvoracle as o

def step(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[{name}] OK {time.time()-t:.2f}s {r if r is not None else ''}", flush=True)
    except Exception:
        print(f"[{name}] FAIL {time.time()-t:.2f}s", flush=True)
        traceback.print_exc(); sys.stdout.flush()

rng = np.random.default_rng(1)
ctx = api.Context(0)
ctxe = api.Context(0, api.FORCE_EXACT_MATCH)
q, t, _ = synth.planted_pair(rng, 300, 700)
oi, od = o.knn2(q, t, 8); oa = o.ratio(oi, od)

def cmp_match(c, q, t, oi, od, oa):
    i, d, a = c.match_knn2(q, t)
    bad = np.nonzero((i != oi).any(1))[0]
    msg = f"path={c.last_match_path} idx_mismatch_rows={len(bad)} dist_eq={np.array_equal(d, od)} acc_eq={np.array_equal(a, oa)}"
    if len(bad):
        r = bad[0]; msg += f" first_bad_row={r} got={i[r]},{d[r]} want={oi[r]},{od[r]}"
    return msg

step("exact int 300x700", lambda: cmp_match(ctxe, q, t, oi, od, oa))
qf = synth.float_desc(rng, 200, 128); tf = synth.float_desc(rng, 500, 128)
fi, fd = o.knn2(qf, tf, 8)
step("exact float 200x500", lambda: cmp_match(ctxe, qf, tf, fi, fd, o.ratio(fi, fd)))
step("tensor int 300x700", lambda: cmp_match(ctx, q, t, oi, od, oa))
q2, t2, _ = synth.planted_pair(rng, 1000, 4000)
i2, d2 = o.knn2(q2, t2, 8)
step("tensor int 1000x4000", lambda: cmp_match(ctx, q2, t2, i2, d2, o.ratio(i2, d2)))

def fh():
    out = []
    for n, rho, dup in [(20, 0.6, 0), (60, 0.5, 0.3), (400, 0.3, 0), (4, 1.0, 0), (8192, 0.3, 0)]:
        s, d, _ = synth.correspondences(rng, n, rho, dup=dup)
        ref = o.find_homography(s, d)
        H, m, rm = ctx.find_homography(s, d, want_ransac_mask=True)
        rel = -1 if H is None or not ref["found"] else float(np.max(np.abs(H - ref["H"]) / np.maximum(np.abs(ref["H"]), 1e-12)))
        out.append((n, "found", H is not None, ref["found"], "rmask_eq", np.array_equal(rm, ref["ransac_mask"]),
                    "mask_diff", int((m != ref["mask"]).sum()), "Hrel", rel, "exact", H is not None and np.array_equal(H, ref["H"])))
    return out
step("find_homography", fh)

def timing():
    rs = []
    q, t, _ = synth.planted_pair(rng, 8192, 8192)
    for c, name in ((ctx, "tensor"), (ctxe, "exact")):
        c.match_knn2(q, t)
        t0 = time.time(); c.match_knn2(q, t); rs.append((name, "8k x 8k e2e ms", round((time.time() - t0) * 1e3, 2)))
    s, d, _ = synth.correspondences(rng, 8192, 0.3)
    ctx.find_homography(s, d)
    t0 = time.time(); ctx.find_homography(s, d); rs.append(("fh n=8192 ms", round((time.time() - t0) * 1e3, 2)))
    s, d, _ = synth.correspondences(rng, 32, 0.5)
    ctx.find_homography(s, d)
    t0 = time.time(); ctx.find_homography(s, d); rs.append(("fh n=32 ms", round((time.time() - t0) * 1e3, 2)))
    return rs
step("timing", timing)
