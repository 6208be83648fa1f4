"""torchrun entry: train-tile sharded kNN over NCCL, checked against the CPU oracle on rank 0."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from computervision_objectdetection_featurematching_b200 import api, sharding, synth

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nq, nt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 65536)
rng = np.random.default_rng(5)                      # same data on every rank
q, t, _ = synth.planted_pair(rng, nq, nt)
a, b = sharding.train_tile_ranges(nt, world)[rank]
ctx = api.Context(local)
qd = torch.from_numpy(q).cuda(); td = torch.from_numpy(t[a:b]).cuda()
idx, dd, acc = sharding.sharded_match(ctx, qd, td, a)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
dist.barrier(); e0.record()
for _ in range(5):
    idx, dd, acc = sharding.sharded_match(ctx, qd, td, a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if rank == 0:
    from oracle import cvoracle as o
    n_chk = min(nq, 512)
    oi, od = o.knn2(q[:n_chk], t, nthreads=os.cpu_count())
    assert np.array_equal(idx.cpu().numpy()[:n_chk], oi) and np.array_equal(dd.cpu().numpy()[:n_chk], od)
    assert np.array_equal(acc.cpu().numpy()[:n_chk], o.ratio(oi, od))
    print(f"sharded match OK world={world} nq={nq} nt={nt} ms={ms:.3f} TFLOP/s={2.0 * nq * nt * 128 / ms / 1e9:.1f}")
dist.barrier()
ctx.close()
dist.destroy_process_group()
