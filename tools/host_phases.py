#!/usr/bin/env python
"""Phase timing of ShardedAllPairs.step_host under torchrun (wall clock around synchronised phases, per rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/host_phases.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fununifrac_b200 import engine, sharded  # noqa: E402


def main():
    world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    N = int(os.environ.get("FUF_N", "10000"))
    inp, leaves = bench.build_tree()
    n = len(inp.basis)
    ctx = engine.TreeContext(inp.parent, inp.edge_length, device=lr)
    job = sharded.ShardedAllPairs(sharded.CudaOps(ctx), N, engine.MODE_L1, world, rank, dist)
    X_pin = torch.zeros((max(job.hi - job.lo, 1), n), dtype=torch.float64, pin_memory=True)
    bench.make_profiles(leaves, n, job.lo, job.hi, X_pin.numpy()[:job.hi - job.lo])
    X_pin = X_pin[:job.hi - job.lo]
    shared = sharded.SharedHostMatrix(N, world, rank, dist)
    stage = torch.empty_like(X_pin, device=f"cuda:{lr}")

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        sync()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dist.barrier()
        torch.cuda.synchronize()
        return (t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3

    for it in range(3):
        job.step_host(X_pin, shared)
    res = {}
    D_rows = torch.empty((max(1, N // world), N), dtype=torch.float64, device=f"cuda:{lr}")   # this rank's share of the matrix
    D_pin = torch.empty(D_rows.shape, dtype=torch.float64, pin_memory=True)
    for it in range(3):
        res["h2d of the shard"] = timed(lambda: stage.copy_(X_pin, non_blocking=True))
        res["step (device)"] = timed(lambda: job.step(stage))
        res["d2h of 1/G of D"] = timed(lambda: D_pin.copy_(D_rows, non_blocking=True))
        res["step_host"] = timed(lambda: job.step_host(X_pin, shared))
    out = torch.tensor([[a, b] for a, b in res.values()], dtype=torch.float64, device=f"cuda:{lr}")
    allr = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(allr, out)
    if rank == 0:
        for i, k in enumerate(res):
            print(f"{k:18s} own ms per rank: " + " ".join(f"{float(a[i, 0]):7.2f}" for a in allr) +
                  f"   | to barrier: {max(float(a[i, 1]) for a in allr):7.2f}", flush=True)
    shared.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
