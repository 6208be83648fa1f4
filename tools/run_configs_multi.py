"""torchrun entry for BASELINE configs 4 and 5 at their full shapes on N GPUs of one box (evidence run, not bench.py).

  C4  4096 image pairs of 4096 x 4096 descriptors, sharded by pair (rank r takes pairs p = r mod N), no data-path
      collective; every pair = kNN + ratio + 2000-iteration RANSAC homography.  Checked: the first pairs of rank 0
      against the CPU oracle (match count, H found, inlier count, RANSAC iterations).
  C5  262 144 queries x 1 048 576 train rows, train-tile sharded (131 072 rows per tile at N = 8), ONE exchange
      (all-gather of 16 B per query and rank over NCCL) + merge; 512 sampled queries checked against the CPU oracle
      over the whole train set.  Then 2^20 RANSAC hypotheses on 8192 correspondences at 30 % inliers (no early stop),
      every rank an independent replica.

usage: torchrun --nproc-per-node N tools/run_configs_multi.py [--scale S]   (S < 1 shrinks every size, for smoke runs)
"""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, sharding, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--skip", default="")
args = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
ctx = api.Context(local)
out = {"n_gpus": world, "scale": args.scale}


def all_max(ms):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ C4
if "c4" not in args.skip:
    N4 = max(256, int(4096 * args.scale) // 64 * 64)
    n_pairs = max(world * 64, int(4096 * args.scale) // (world * 64) * (world * 64))
    B = 64
    POOL = 4                                                       # distinct scene batches per rank (> L2 in bf16)
    rngq = np.random.default_rng(4000)
    q = synth.sift_like(rngq, N4); qk = rngq.uniform([0, 0], [640, 480], size=(N4, 2)).astype(np.float32)
    models = ctx.upload_models(q, qk, [0, N4], [0])

    def make_pair(p):
        """Scene p: planted copies of half the view's descriptors, 30 % of them on a homography."""
        rng = np.random.default_rng(4000 + 1 + (p % (POOL * B * world)))
        t = synth.sift_like(rng, N4); tk = rng.uniform([0, 0], [640, 480], size=(N4, 2)).astype(np.float32)
        k = N4 // 2
        rq = rng.permutation(N4)[:k]; rt = rng.permutation(N4)[:k]
        t[rt] = np.clip(q[rq] + np.round(rng.normal(0, 12.0, size=(k, 128))).astype(np.float32), 0, 255)
        H = synth.random_homography(rng)
        geo = rng.random(k) < 0.3
        pp = np.c_[qk[rq[geo]], np.ones(int(geo.sum()))] @ H.T
        tk[rt[geo]] = (pp[:, :2] / pp[:, 2:3] + rng.normal(0, 0.7, size=(int(geo.sum()), 2))).astype(np.float32)
        return t, tk

    mine = list(range(rank, n_pairs, world))                      # pair p -> GPU p mod N
    pool = []
    for b in range(POOL):
        ps = mine[b * B:(b + 1) * B]
        ts, ks = zip(*[make_pair(p) for p in ps])
        pool.append(ctx.upload_scenes(np.concatenate(ts), np.concatenate(ks), np.arange(B + 1, dtype=np.int64) * N4))
    steps = len(mine) // B
    ctx.detect_scenes(models, pool[0])                             # warm-up
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0.record(stream)
    crc = 0; accepted = 0; first_res = None
    for s in range(steps):
        res = ctx.detect_scenes(models, pool[s % POOL])
        if s == 0:
            first_res = res.copy()
        if s < POOL:                                               # checksum of the distinct pairs only
            crc ^= zlib.crc32(np.ascontiguousarray(res["status"]).tobytes() + np.ascontiguousarray(res["n_inliers"]).tobytes()
                              + np.ascontiguousarray(res["ransac_iters"]).tobytes())
        accepted += int((res["status"] == 0).sum())
    e1.record(stream); e1.synchronize()
    ms = all_max(e0.elapsed_time(e1))
    t = torch.tensor([crc, accepted], device=dev, dtype=torch.int64)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    out["c4"] = {"pairs": steps * B * world, "nq": N4, "nt": N4, "ms": ms, "pairs_per_s": steps * B * world / (ms * 1e-3),
                 "match_tflops_equiv": 2.0 * N4 * N4 * 128 * steps * B * world / (ms * 1e-3) / 1e12,
                 "accepted": int(sum(int(a[1]) for a in allc)),
                 "per_rank_crc": [int(a[0]) for a in allc]}
    if rank == 0:
        from oracle import cvoracle as o
        ok = True
        for j in range(2):
            t, tk = make_pair(mine[j])
            oi, od = o.knn2(q, t, nthreads=os.cpu_count()); oa = o.ratio(oi, od).astype(bool)
            ref = o.find_homography(qk[oa], tk[oi[oa, 0]])
            r = first_res[j, 0]
            ok = ok and int(r["n_good"]) == int(oa.sum()) and (int(r["status"]) != api.H_EMPTY) == ref["found"] \
                and int(r["ransac_iters"]) == ref["info"]["iters_run"] \
                and (not ref["found"] or abs(int(r["n_inliers"]) - int(ref["mask"].sum())) <= 2)
        out["c4"]["oracle_check_first_pairs"] = bool(ok)
    for sc in pool:
        sc.free()
    models.free()

# ------------------------------------------------------------------------------------------------ C5
if "c5" not in args.skip:
    NQ = max(1024, int(262144 * args.scale) // 128 * 128)
    NT = max(world * 1024, int(1048576 * args.scale) // (world * 256) * (world * 256))
    ranges = sharding.train_tile_ranges(NT, world)
    a, b = ranges[rank]
    q = synth.sift_like(np.random.default_rng(5000), NQ)           # replicated queries (same seed everywhere)

    def make_tile(r):
        ra, rb = ranges[r]
        rng = np.random.default_rng(5001 + r)
        t = synth.sift_like(rng, rb - ra)
        k = min(NQ // world // 2, rb - ra)                          # planted noisy copies of some queries in every tile
        rq = rng.permutation(NQ)[:k]; rt = rng.permutation(rb - ra)[:k]
        t[rt] = np.clip(q[rq] + np.round(rng.normal(0, 12.0, size=(k, 128))).astype(np.float32), 0, 255)
        return t

    t_mine = make_tile(rank)
    qd = torch.from_numpy(q).to(dev); td = torch.from_numpy(t_mine).to(dev)
    idx, dd, acc = sharding.sharded_match(ctx, qd, td, a)          # warm-up (allocations, NCCL setup)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, dd, acc = sharding.sharded_match(ctx, qd, td, a)
    e1.record(); e1.synchronize()
    ms = all_max(e0.elapsed_time(e1))
    c5 = {"nq": NQ, "nt": NT, "tile_rows": b - a, "ms": ms, "match_tflops": 2.0 * NQ * NT * 128 / (ms * 1e-3) / 1e12,
          "accepted": int(acc.sum().item())}
    if rank == 0:
        from oracle import cvoracle as o
        t0 = time.time()
        t_all = np.concatenate([t_mine] + [make_tile(r) for r in range(1, world)])
        n_chk = 512
        rows = np.random.default_rng(5999).permutation(NQ)[:n_chk]
        oi, od = o.knn2(q[rows], t_all, nthreads=os.cpu_count())
        gi = idx.cpu().numpy()[rows]; gd = dd.cpu().numpy()[rows]; ga = acc.cpu().numpy()[rows]
        c5["checked_queries"] = n_chk
        c5["idx_equal"] = bool(np.array_equal(gi, oi)); c5["dist_equal"] = bool(np.array_equal(gd, od))
        c5["accept_equal"] = bool(np.array_equal(ga, o.ratio(oi, od)))
        c5["oracle_s"] = time.time() - t0
        del t_all
    out["c5_match"] = c5
    del qd, td
    torch.cuda.empty_cache()

    # 2^20 hypotheses on 8192 correspondences at 30 % inliers, scored exhaustively (no early stop, no refit)
    n_hyp = max(4096, int((1 << 20) * args.scale))
    src, dst, _ = synth.correspondences(np.random.default_rng(5002), 8192, 0.3)
    H, mask = ctx.find_homography(src, dst, max_iters=n_hyp, flags=api.RANSAC_NO_EARLY_STOP | api.RANSAC_NO_REFINE)
    barrier()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    H, mask = ctx.find_homography(src, dst, max_iters=n_hyp, flags=api.RANSAC_NO_EARLY_STOP | api.RANSAC_NO_REFINE)
    e1.record(stream); e1.synchronize()
    ms = all_max(e0.elapsed_time(e1))
    c5r = {"hypotheses": n_hyp, "n": 8192, "ms_incl_h2d_d2h": ms, "hyps_per_s_per_gpu": n_hyp / (ms * 1e-3),
           "algorithmic_gbs_per_gpu": 16.0 * 8192 * n_hyp / (ms * 1e-3) / 1e9, "best_count": int(mask.sum())}
    if rank == 0:
        from oracle import cvoracle as o
        ref = o.ransac_stage(src, dst, max_iters=2000)
        c5r["oracle_best_count_2000_iters"] = int(ref["mask"].sum())
        c5r["count_not_worse"] = bool(mask.sum() >= ref["mask"].sum())
    out["c5_ransac"] = c5r

if rank == 0:
    print(json.dumps(out), flush=True)
barrier()
ctx.close()
dist.destroy_process_group()
