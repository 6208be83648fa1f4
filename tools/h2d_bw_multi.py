"""Raw host -> device bandwidth of N GPUs copying at the same time (isolates the host side of bench.py's e2e scaling):
every rank copies a 272.6 MB pinned buffer (one e2e step of fp32 descriptors) back to back, (a) alone on an otherwise idle
box (rank by rank), (b) all ranks at once, (c) all ranks at once while every GPU also runs the resident c3 workload.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bw_multi.py
Prints one JSON line on rank 0."""
import json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from computervision_objectdetection_featurematching_b200 import api

local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
numa = bench.bind_to_gpu_numa_node(torch.cuda.get_device_properties(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 272630280
host = torch.empty(N, dtype=torch.uint8).pin_memory(); host.fill_(1)
devb = torch.empty(N, dtype=torch.uint8, device="cuda")
copy_stream = torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def copies(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        e0.record()
        for _ in range(reps):
            devb.copy_(host, non_blocking=True)
        e1.record()
    e1.synchronize()
    return reps * N / (e0.elapsed_time(e1) * 1e-3) / 1e9


def gather(v):
    if world == 1:
        return [v]
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


copies(3)
alone = []
for r in range(world):                       # (a) one rank at a time
    barrier()
    v = copies(20) if r == rank else 0.0
    barrier()
    alone.append(max(gather(v)))
barrier()
together = gather(copies(20))                # (b) all at once
# (c) all at once with the GPUs busy: the resident c3 workload on one context per GPU, fed by a second thread
q, qk, batches = bench.make_workload(3000 + rank, 64, 1)
ctx = api.Context(local)
models = ctx.upload_models(q, qk, [0, bench.NQ], [0])
sc = ctx.upload_scenes(*batches[0])
import threading
stop = False; steps = 0


def compute():
    global steps
    torch.cuda.set_device(local)
    jobs = []
    while not stop:
        jobs.append(ctx.submit_scenes(models, sc, want_inliers=False))
        if len(jobs) > 2:
            jobs.pop(0).wait(); steps += 1
    for j in jobs:
        j.wait()


th = threading.Thread(target=compute); th.start()
time.sleep(0.3)
barrier()
t0 = time.time(); s0 = steps
busy = gather(copies(20))
rate = (steps - s0) * 64 / (time.time() - t0)
stop = True; th.join()
rates = gather(rate)
if rank == 0:
    print(json.dumps({"n_gpus": world, "buffer_bytes": N, "host_numa_node_rank0": numa, "cpus": os.cpu_count(),
                      "h2d_gbs_alone_per_gpu": alone, "h2d_gbs_all_at_once_per_gpu": together, "h2d_gbs_all_at_once_sum": sum(together),
                      "h2d_gbs_all_at_once_gpus_busy_per_gpu": busy, "h2d_gbs_all_at_once_gpus_busy_sum": sum(busy),
                      "resident_pairs_per_s_while_copying_per_gpu": rates}), flush=True)
sc.free(); models.free(); ctx.close()
if world > 1:
    barrier(); dist.destroy_process_group()
