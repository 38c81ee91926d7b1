#!/usr/bin/env python
"""Phase timing (CUDA events at the marks of ShardedAllPairs.step) of the device-resident multi-GPU step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/step_phases.py [l1|l2] [N]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fununifrac_b200 import engine, sharded  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "l1"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else (10000 if what == "l1" else 50000)
    world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    inp, leaves = bench.build_tree()
    ctx = engine.TreeContext(inp.parent, inp.edge_length, device=lr)
    mode = engine.MODE_L1 if what == "l1" else engine.MODE_L2
    job = sharded.ShardedAllPairs(sharded.CudaOps(ctx), N, mode, world, rank, dist)
    X = bench.device_profiles(leaves, len(inp.basis), job.hi - job.lo, seed=1 + rank)
    for _ in range(3):
        job.step(X)
    order = ["start", "push_begin", "push_end", "pair_begin", "pair_end", "refine_end", "gather_end", "end"]
    reps, acc, names = 5, None, None
    for _ in range(reps):
        ev = {}

        def mark(name):
            ev[name] = torch.cuda.Event(enable_timing=True)
            ev[name].record()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        job.step(X, mark)
        torch.cuda.synchronize()
        names = [n for n in order if n in ev]          # (the peer form has no gather phase)
        t = torch.tensor([ev[a].elapsed_time(ev[b]) for a, b in zip(names, names[1:])], dtype=torch.float64)
        acc = t if acc is None else acc + t
    acc /= reps
    out = [torch.zeros_like(acc).cuda() for _ in range(world)]
    dist.all_gather(out, acc.cuda())
    if rank == 0:
        print(f"{what} N={N} world={world} peer={job.use_peer}: phase ms per rank: " + " | ".join(f"{a}->{b}" for a, b in zip(names, names[1:])))
        for r, t in enumerate(out):
            print(f"  rank {r}: " + "  ".join(f"{float(v):7.3f}" for v in t) + f"   sum {float(t.sum()):7.3f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
