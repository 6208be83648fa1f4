"""Train-tile sharding on the GPU.  With one GPU the shards run one after another on the same device
(all shards' data in one process, as B200_PROFILING.md asks for when ranks outnumber GPUs); with >= 2
GPUs visible the same path also runs as 2 NCCL ranks under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

from computervision_objectdetection_featurematching_b200 import sharding, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_logical_shards_on_one_gpu(oracle):
    import torch
    from computervision_objectdetection_featurematching_b200 import api
    rng = np.random.default_rng(21)
    datasets = []
    q, t, _ = synth.planted_pair(rng, 700, 5000)
    t[2600] = t[100]; q[9] = t[100]                              # tie across shards
    datasets.append((q, t))
    qf = synth.float_desc(rng, 700, 128); tf = synth.float_desc(rng, 5000, 128)      # candidate + re-rank path per shard
    tf[3100] = tf[40]; qf[5] = tf[40]; tf[4000:4100] = tf[:100] + np.float32(1e-3)
    datasets.append((qf, tf))
    dev = torch.device("cuda", 0)
    with api.Context(0) as ctx:
      for q, t in datasets:
        oi, od = oracle.knn2(q, t, nthreads=8)
        oa = oracle.ratio(oi, od)
        qd = torch.from_numpy(q).to(dev); td = torch.from_numpy(t).to(dev)
        for world in (1, 2, 4):
            dp = torch.empty((world, 700, 2), dtype=torch.float32, device=dev)
            ip = torch.empty((world, 700, 2), dtype=torch.int32, device=dev)
            for r, (a, b) in enumerate(sharding.train_tile_ranges(5000, world)):
                ctx.dev_match_top2(qd.data_ptr(), 700, td[a:b].data_ptr(), b - a, a, dp[r].data_ptr(), ip[r].data_ptr(),
                                   stream=torch.cuda.current_stream().cuda_stream)
            idx = torch.empty((700, 2), dtype=torch.int32, device=dev)
            dd = torch.empty((700, 2), dtype=torch.float32, device=dev)
            acc = torch.empty((700,), dtype=torch.uint8, device=dev)
            ctx.dev_merge_top2(dp.data_ptr(), ip.data_ptr(), world, 700, 0.9, idx.data_ptr(), dd.data_ptr(), acc.data_ptr(),
                               stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(idx.cpu().numpy(), oi), world
            assert np.array_equal(dd.cpu().numpy(), od), world
            assert np.array_equal(acc.cpu().numpy(), oa), world


def test_two_nccl_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tools", "run_sharded_match.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sharded match OK" in r.stdout
