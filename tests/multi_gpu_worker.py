"""Launched by tests/test_gpu_parity.py under torchrun (one process per GPU): the sharded all-pairs job through
``sharded.ShardedAllPairs`` on real GPUs — peer form (shard pairs, NVLink pulls, in-place stores into rank 0's matrix or
the shared host matrix) for weighted L1, the direct --L2 kernel and the tensor-core Gram path — against the CPU oracle.

    torchrun --nproc-per-node G tests/multi_gpu_worker.py N mode        (mode: l1 | l2)
Prints ``MULTI_GPU_OK`` on rank 0 when every check passed."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fununifrac_b200 import engine, sharded, synth  # noqa: E402
from fununifrac_b200.factory import make_emd_input, make_tree  # noqa: E402
from oracle import oracle_c  # noqa: E402


class HostStagedDist:
    """``torch.distributed`` over gloo for CUDA tensors, staged through host memory — lets several ranks share ONE GPU
    (NCCL refuses that), so that the peer form's CUDA mechanics (IPC mappings, copy-engine pulls, arrival flags, in-place
    stores) can be exercised on a single-GPU box.  Only the calls sharded.py makes."""

    ReduceOp = dist.ReduceOp

    @staticmethod
    def _sync():
        torch.cuda.synchronize()

    def broadcast(self, t, src):
        self._sync()
        h = t.cpu()
        dist.broadcast(h, src=src)
        t.copy_(h)

    def all_reduce(self, t, op=dist.ReduceOp.SUM):
        self._sync()
        h = t.cpu()
        dist.all_reduce(h, op=op)
        t.copy_(h)

    def all_gather_into_tensor(self, out, inp):
        self._sync()
        h = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(h, inp.cpu().contiguous())
        out.copy_(h)

    def gather(self, t, gather_list=None, dst=0):
        self._sync()
        hs = [torch.empty(t.shape, dtype=t.dtype) for _ in gather_list] if gather_list is not None else None
        dist.gather(t.cpu(), hs, dst=dst)
        if gather_list is not None:
            for g, h in zip(gather_list, hs):
                g.copy_(h)

    def all_gather_object(self, out, obj):
        dist.all_gather_object(out, obj)

    def broadcast_object_list(self, objs, src=0):
        dist.broadcast_object_list(objs, src=src)

    def barrier(self):
        self._sync()
        dist.barrier()


def main():
    N, l2 = int(sys.argv[1]), sys.argv[2] == "l2"
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    one_gpu = torch.cuda.device_count() < world          # several ranks on device 0: gloo staged through the host
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group("gloo")
        comm = HostStagedDist()
    else:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        comm = dist
    with tempfile.TemporaryDirectory() as tmp:
        path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
        inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
    ctx = engine.TreeContext(inp.parent, inp.edge_length, device=local)
    n = ctx.n
    leaves = synth.leaves_of(inp.parent)
    X = synth.synth_profiles_dense(leaves, n, N, seed=41)                # the same on every rank
    rng = np.random.default_rng(41)
    X[N - 2] = X[3] * (1 + 1e-6 * rng.random(n))                         # near-duplicate pair across the first and last shard
    X[N - 2] /= X[N - 2].sum()
    X[N // 2 + 1] = X[7]                                                 # identical pair
    mode = engine.MODE_L2 if l2 else engine.MODE_L1
    ops = sharded.CudaOps(ctx)
    if os.environ.get("FUF_TEST_VERBOSE") == "2":        # debugging aid: synchronise after every ops call, name the one that fails
        class Traced:
            def __init__(self, inner):
                self.__dict__["_inner"] = inner

            def __getattr__(self, name):
                attr = getattr(self._inner, name)
                if not callable(attr):
                    return attr

                def call(*a, **k):
                    out = attr(*a, **k)
                    try:
                        torch.cuda.synchronize()
                    except Exception:
                        print(f"rank {rank}: CUDA error surfaced after ops.{name}", flush=True)
                        raise
                    return out
                return call
        ops = Traced(ops)
    job = sharded.ShardedAllPairs(ops, N, mode, world, rank, comm)
    assert job.use_peer and job.use_gram == (l2 and N >= 512)
    Xl = engine.profile_matrix(job.hi - job.lo, n, torch.float64, f"cuda:{local}")
    Xl.copy_(torch.from_numpy(X[job.lo:job.hi]))
    results = []
    for it in range(3):                                                  # epochs advance, buffers are reused
        D = job.step(Xl)
        torch.cuda.synchronize()
        if os.environ.get("FUF_TEST_VERBOSE"):
            print(f"rank {rank}: step {it} done", flush=True)
        if rank == 0:
            results.append(D.cpu().numpy().copy())
    shared = sharded.SharedHostMatrix(N, world, rank, comm)
    Xh = engine.profile_matrix(job.hi - job.lo, n, torch.float64, None, pin_memory=True)
    Xh.copy_(torch.from_numpy(X[job.lo:job.hi]))
    Dh = None
    for _ in range(2):
        if rank == 0:
            shared.array[...] = np.nan
        comm.barrier()
        Dh = job.step_host(Xh, shared)
    if rank == 0:
        ref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, l2), l2)
        iu = np.triu_indices(N, 1)
        mass = ((oracle_c.push_up(inp.parent, inp.edge_length, X[:4], l2) ** (2 if l2 else 1)).sum(1)).max()
        for got in results + [Dh]:
            assert np.array_equal(got, got.T) and not np.diag(got).any() and not np.isnan(got).any()
            ok = ref[iu] > 0
            err = np.abs(got[iu][ok] - ref[iu][ok]) / ref[iu][ok]
            floor = np.abs(got[iu][ok] - ref[iu][ok]) <= 64 * np.finfo(np.float64).eps * mass
            assert np.all((err < 1e-6) | floor), float(err[~floor].max())
            assert got[7, N // 2 + 1] == 0.0
        assert all(np.array_equal(results[0], r) for r in results[1:])    # deterministic from step to step
        assert np.allclose(Dh, results[0], rtol=1e-12, atol=0)
        assert job.last_flagged >= 1                                      # the near-duplicate went through the list path
        print(f"MULTI_GPU_OK world={world} N={N} mode={'l2' if l2 else 'l1'} gram={job.use_gram} flagged={job.last_flagged}",
              flush=True)
    comm.barrier()
    shared.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
