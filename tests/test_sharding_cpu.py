"""Host-side logic of the N>1 paths on CPU: world_size-2 gloo processes exchange per-shard top-2 results
(computed here by the oracle standing in for the GPU kernel) and merge them; the result must equal the
unsharded oracle result, ties included.  Also the cost-balanced pair sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from computervision_objectdetection_featurematching_b200 import sharding, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, t, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cvoracle as o
    a, b = sharding.train_tile_ranges(t.shape[0], world, align=256)[rank]
    idx, dd = o.knn2(q, t[a:b])
    idx = np.where(idx >= 0, idx + a, -1).astype(np.int32)
    dd = np.where(idx >= 0, dd, np.inf).astype(np.float32)
    d_all = [torch.empty_like(torch.from_numpy(dd)) for _ in range(world)]
    i_all = [torch.empty_like(torch.from_numpy(idx)) for _ in range(world)]
    dist.all_gather(d_all, torch.from_numpy(dd)); dist.all_gather(i_all, torch.from_numpy(idx))
    mi, md, ma = sharding.merge_top2_reference(torch.stack(d_all).numpy(), torch.stack(i_all).numpy())
    if rank == 0:
        out.put((mi, md, ma))
    dist.barrier()
    dist.destroy_process_group()


def test_train_tile_sharding_gloo_world2(oracle):
    rng = np.random.default_rng(11)
    q, t, _ = synth.planted_pair(rng, 200, 900)
    t[300] = t[10]; t[700] = t[10]; q[5] = t[10]            # exact ties straddling the shard boundary
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, t, out)) for r in range(2)]
    for p in procs:
        p.start()
    mi, md, ma = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(mi, oi) and np.array_equal(md, od)
    assert np.array_equal(ma, oracle.ratio(oi, od))


def test_merge_is_shard_count_invariant(oracle):
    rng = np.random.default_rng(12)
    q, t, _ = synth.planted_pair(rng, 64, 1500)
    oi, od = oracle.knn2(q, t)
    for world in (1, 2, 3, 5, 8):
        dp, ip = [], []
        for a, b in sharding.train_tile_ranges(1500, world, align=256):
            i, d = oracle.knn2(q, t[a:b])
            ip.append(np.where(i >= 0, i + a, -1)); dp.append(np.where(i >= 0, d, np.inf))
        mi, md, _ = sharding.merge_top2_reference(np.stack(dp).astype(np.float32), np.stack(ip).astype(np.int32))
        assert np.array_equal(mi, oi) and np.array_equal(md, od), world


def test_pair_sharding_balances_cost():
    rng = np.random.default_rng(13)
    costs = rng.integers(200, 9000, size=150).astype(float)
    for world in (2, 4, 8):
        parts = sharding.assign_by_cost(costs, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(150))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.05
    ranges = sharding.train_tile_ranges(1 << 20, 8)
    assert ranges[0] == (0, 131072) and ranges[-1][1] == 1 << 20 and all(a % 256 == 0 for a, _ in ranges)
