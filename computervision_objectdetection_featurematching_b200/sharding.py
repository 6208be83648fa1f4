"""Multi-GPU decomposition of the hot path.  One process per GPU; torch.distributed is
the plumbing (NCCL on GPUs, gloo in the CPU tests).

1. Pair sharding — (model view x test image) pairs are independent (reference src/TestsDetector.cpp:58-95
   carries no state between iterations), so scenes are dealt to ranks by cost and every rank runs the
   fused path on its own scenes against the replicated model set.  No data-path collective.
2. Train-tile sharding — one huge train set is cut into contiguous row ranges, each rank computes its
   local top-2 per query (cvg_dev_match_top2), the partial results are all-gathered (16 B per query and
   rank) and merged in lexicographic (distance, global train index) order (cvg_dev_merge_top2), which
   reproduces cv::BFMatcher's tie rule independently of the shard count.
"""
import numpy as np


def assign_by_cost(costs, world):
    """Longest-processing-time-first assignment of items (scenes) to `world` ranks.
    costs[i] ~ Nq * Nt_i.  Returns a list of index arrays, one per rank (each sorted ascending)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    out = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(int(i)); load[r] += costs[i]
    return [np.array(sorted(x), dtype=np.int64) for x in out]


def train_tile_ranges(n_train, world, align=256):
    """Contiguous train-row ranges per rank, boundaries aligned to the 256-row MMA tile."""
    per = -(-n_train // world)
    per = -(-per // align) * align
    return [(min(r * per, n_train), min((r + 1) * per, n_train)) for r in range(world)]


def merge_top2_reference(dist_parts, idx_parts, ratio=0.9):
    """Host restatement of the merge rule, [P, nq, 2] partial results -> idx, dist, accept.
    Used by the CPU tests of the N>1 logic; the product path merges on the GPU (cvg_dev_merge_top2)."""
    P, nq, _ = dist_parts.shape
    d = np.transpose(dist_parts, (1, 0, 2)).reshape(nq, 2 * P).astype(np.float32)
    i = np.transpose(idx_parts, (1, 0, 2)).reshape(nq, 2 * P).astype(np.int64)
    dk = np.where(i >= 0, d, np.float32(np.inf))
    ik = np.where(i >= 0, i, np.iinfo(np.int64).max)
    order = np.lexsort((ik, dk), axis=1)[:, :2]                      # primary: distance, secondary: index
    rows = np.arange(nq)[:, None]
    idx = np.where(np.isfinite(dk[rows, order]) | (i[rows, order] >= 0), i[rows, order], -1)
    idx = np.where(ik[rows, order] == np.iinfo(np.int64).max, -1, idx).astype(np.int32)
    dist = np.where(idx >= 0, d[rows, order], 0).astype(np.float32)
    accept = ((idx[:, 0] >= 0) & (idx[:, 1] >= 0) & (dist[:, 0] < np.float32(ratio) * dist[:, 1])).astype(np.uint8)
    return idx, dist, accept


def sharded_match(ctx, query_dev, train_tile_dev, index_base, ratio=0.9, group=None):
    """Train-tile sharded kNN on GPUs.  query_dev: [nq,128] float32 CUDA tensor replicated on every rank;
    train_tile_dev: this rank's [nt_r,128] CUDA tensor; index_base: global row of its first train row.
    Returns CUDA tensors (idx [nq,2] int32, dist [nq,2] float32, accept [nq] uint8), identical on all ranks."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nq = query_dev.shape[0]
    dev = query_dev.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    d_loc = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    i_loc = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    ctx.dev_match_top2(query_dev.data_ptr(), nq, train_tile_dev.data_ptr(), train_tile_dev.shape[0], int(index_base),
                       d_loc.data_ptr(), i_loc.data_ptr(), stream=stream)
    if world > 1:
        d_all = torch.empty((world, nq, 2), dtype=torch.float32, device=dev)
        i_all = torch.empty((world, nq, 2), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(d_all, d_loc, group=group)        # the one exchange step of the path
        dist.all_gather_into_tensor(i_all, i_loc, group=group)
    else:
        d_all, i_all = d_loc[None], i_loc[None]
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    dd = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    acc = torch.empty((nq,), dtype=torch.uint8, device=dev)
    ctx.dev_merge_top2(d_all.data_ptr(), i_all.data_ptr(), world, nq, ratio, idx.data_ptr(), dd.data_ptr(),
                       acc.data_ptr(), stream=stream)
    return idx, dd, acc
