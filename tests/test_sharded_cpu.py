"""Host-side logic of the multi-GPU path (fununifrac_b200/sharded.py) on CPU: the partition helpers, and the
whole exchange sequence with world_size 2 and 3 over the ``gloo`` backend.  The compute calls are replaced by a
test-only stand-in built on the CPU oracle (the product always injects the CUDA context); what is under test is
the sharding, the layouts that cross ranks, and the assembly of the full matrix on rank 0."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fununifrac_b200 import sharded

TILE = 128


def test_partition_helpers():
    for N, world in ((1, 1), (127, 2), (128, 2), (700, 3), (10000, 8), (10000, 1), (300, 8)):
        sh = sharded.shard_size(N, world)
        assert sh % TILE == 0 and sh * world >= N
        ranges = [sharded.sample_range(N, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        T = (N + TILE - 1) // TILE
        rows = [sharded.owned_tile_rows(N, world, r) for r in range(world)]
        assert sorted(sum(rows, [])) == list(range(T))
        tiles = [sharded.tiles_of(r, N) for r in rows]
        assert sum(tiles) == T * (T + 1) // 2
        if T >= 4 * world:          # balanced to within ~one tile row
            assert max(tiles) - min(tiles) <= T
    offs, total = sharded.packed_offsets([0, 5, -1, 2], 700)
    assert offs == [0, 128 * 700, 128 * 700 + 128 * 60, 128 * 700 + 128 * 60] and total == offs[-1] + 128 * (700 - 256)
    # 10K samples over 8 GPUs: every rank within 2% of the mean tile count
    t = [sharded.tiles_of(sharded.owned_tile_rows(10000, 8, r), 10000) for r in range(8)]
    assert max(t) / (sum(t) / 8) < 1.02


class OracleOps:
    """Test-only stand-in for ``sharded.CudaOps``: same calls, computed by the CPU oracle on torch CPU tensors."""

    def __init__(self, parent, length, l2):
        from oracle import oracle_c
        self.o, self.parent, self.length, self.l2, self.n = oracle_c, parent, length, l2, len(parent)
        self.refine_calls = 0

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=getattr(torch, dtype))

    def push_up_center(self, X, mode, out):
        m = min(X.shape[0], 32)
        c = self.o.push_up(self.parent, self.length, X[:m].numpy().mean(0), self.l2)
        has_child = np.zeros(self.n, bool)
        has_child[self.parent[self.parent >= 0]] = True
        out.copy_(torch.from_numpy(np.where(has_child, c, 0.0)))

    def push_up(self, X, mode, PT, col0, center, P64T, err_stat):
        P = self.o.push_up(self.parent, self.length, X.numpy(), self.l2)            # [N, n]
        N = P.shape[0]
        c = P - center.numpy()[None, :]
        f = c.astype(np.float32)
        PT[:, col0:col0 + N] = torch.from_numpy(f.T.copy())
        if P64T is not None:
            P64T[:, col0:col0 + N] = torch.from_numpy(P.T.copy())
        e = np.abs(c - f.astype(np.float64))
        err_stat[col0:col0 + N] = torch.from_numpy((e ** 2).sum(1) if self.l2 else e.sum(1))

    def push_up_p64(self, X, mode, P64T):
        P = self.o.push_up(self.parent, self.length, X.numpy(), self.l2)
        P64T[:, :P.shape[0]] = torch.from_numpy(P.T.copy())
        self.p64_calls = getattr(self, "p64_calls", 0) + 1

    def _samples(self, PS, N, shard):
        G, n, sh = PS.shape
        return PS.permute(0, 2, 1).reshape(G * sh, n)[:N].double().numpy()

    def pairwise(self, PT, N, mode, tile_rows, shard, shard_stride, D):
        P = self._samples(PT, N, shard)
        if tile_rows is None:
            D.copy_(torch.from_numpy(self.o.pairwise(P, self.l2)))
            return D
        D.fill_(float("nan"))       # everything the contract does not promise is poisoned
        for c, t in enumerate(tile_rows):
            r0, r1 = t * TILE, min(N, (t + 1) * TILE)
            blk = self.o.pairwise_rows(P, r0, r1, self.l2)
            D[c * TILE:c * TILE + (r1 - r0), r0:] = torch.from_numpy(blk[:, r0:])
            lower = np.tril_indices(r1 - r0, -1)      # the kernels only promise i <= j inside the diagonal tile
            D[c * TILE + lower[0], r0 + lower[1]] = float("nan")
        return D

    def count_uncertified_into(self, N, mode, err_stat, D, tile_rows, lin_stat, out):
        assert err_stat.numel() >= N
        out[0] = 1 if tile_rows is None or 0 in tile_rows else 0   # pretend the owner of tile row 0 found one pair

    def refine(self, P64T, N, mode, err_stat, D, tile_rows, shard, shard_stride, lin_stat=None):
        self.refine_calls += 1
        assert P64T.shape[0] * P64T.shape[2] >= N and err_stat.numel() >= N
        return D

    def pairwise_f64(self, P64T, N, mode, tile_rows, shard, shard_stride, D):
        self.f64_calls = getattr(self, "f64_calls", 0) + 1
        return self.pairwise(P64T, N, mode, tile_rows, shard, shard_stride, D)

    def assemble(self, blocks, rows, N, D):
        blocks = blocks.reshape(-1, blocks.shape[-1])
        for c, t in enumerate(rows):
            if t < 0:
                continue
            r0, r1 = t * TILE, min(N, (t + 1) * TILE)
            blk = blocks[c * TILE:c * TILE + (r1 - r0)]
            D[r0:, r0:r1] = blk[:, r0:].T
            D[r0:r1, r1:] = blk[:, r1:]
            D[r0:r1, r0:r1] = torch.triu(blk[:, r0:r1]) + torch.triu(blk[:, r0:r1], 1).T
        return D

    def pack_tile_rows(self, blocks, rows, offsets, N, packed):
        packed.fill_(float("nan"))
        for c, (t, off) in enumerate(zip(rows, offsets)):
            r0, r1 = t * TILE, min(N, (t + 1) * TILE)
            packed[off:off + (r1 - r0) * (N - r0)] = blocks[c * TILE:c * TILE + (r1 - r0), r0:].reshape(-1)
        return packed

    def assemble_packed(self, packed, rows, offsets, N, D):
        packed = packed.reshape(-1)
        for t, off in zip(rows, offsets):
            if t < 0:
                continue
            r0, r1 = t * TILE, min(N, (t + 1) * TILE)
            blk = packed[off:off + (r1 - r0) * (N - r0)].reshape(r1 - r0, N - r0)
            D[r0:, r0:r1] = blk.T
            D[r0:r1, r1:] = blk[:, r1 - r0:]
            D[r0:r1, r0:r1] = torch.triu(blk[:, :r1 - r0]) + torch.triu(blk[:, :r1 - r0], 1).T
        return D

    # host-buffer variant (step_host)
    def to_device(self, X_host, stage):
        src = X_host if torch.is_tensor(X_host) else torch.from_numpy(X_host)
        stage[:src.shape[0]].copy_(src)
        return stage[:src.shape[0]]

    def store_tile_rows_host(self, blocks, rows, N, D_host):
        self.assemble(blocks, rows, N, torch.from_numpy(D_host))

    def copy_to_host(self, D, D_host):
        D_host[...] = D.numpy()

    def synchronize(self):
        pass


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, l2, out_path):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        n = 300
        parent = np.minimum(np.arange(n) + 1 + rng.integers(0, 6, n), n - 1).astype(np.int32)
        parent[-1] = -1
        length = rng.random(n) * (rng.random(n) > 0.1)
        X = rng.random((N, n)) * (rng.random((N, n)) < 0.3)
        X[:, -1] = 0
        X /= X.sum(1, keepdims=True)
        ops = OracleOps(parent, length, l2)
        job = sharded.ShardedAllPairs(ops, N, int(l2), world, rank, dist)
        marks = []
        D = job.step(torch.from_numpy(X[job.lo:job.hi]), marks.append)
        assert marks[0] == "start" and marks[-1] == "end" and "pair_begin" in marks
        assert ops.refine_calls == (1 if job.rows else 0) and job.last_flagged == 1
        # the float64 profiles were pushed up only because a pair was flagged, by every rank that holds samples
        assert getattr(ops, "p64_calls", 0) == (1 if job.hi > job.lo else 0)
        if rank == 0:
            from oracle import oracle_c
            ref = oracle_c.pairwise(oracle_c.push_up(parent, length, X, l2), l2)
            got = D.numpy()
            assert not np.isnan(got).any()
            iu = np.triu_indices(N, 1)
            err = np.abs(got[iu] - ref[iu]) / ref[iu]
            assert err.max() < 1e-5, err.max()          # fp32-rounded operands, float64 arithmetic
            assert np.array_equal(got, got.T) and not np.diag(got).any()
            np.save(out_path, got)
        else:
            assert D is None
        # host-buffer variant: every rank stores its rows into one shared host matrix; all ranks see the full result
        shared = sharded.SharedHostMatrix(N, world, rank, dist, register=False)
        shared.array.fill(np.nan)
        dist.barrier()
        Dh = job.step_host(X[job.lo:job.hi], shared)
        want = torch.zeros(1, dtype=torch.float64)
        if rank == 0:
            want[0] = float(np.abs(got).sum())
            assert np.array_equal(Dh, got)
        dist.broadcast(want, src=0)
        assert not np.isnan(Dh).any() and float(np.abs(Dh).sum()) == float(want[0])      # same pages on every rank
        dist.barrier()
        del Dh
        shared.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,l2", [(2, 300, False), (3, 700, False), (2, 129, True), (3, 100, False)])
def test_sharded_exchange_gloo(world, N, l2, tmp_path, monkeypatch):
    if world == 3:      # once through the named shared-memory file instead of /proc/<pid>/fd
        monkeypatch.setenv("FUF_SHARED_NO_PROCFD", "1")
    out = str(tmp_path / "D.npy")
    mp.spawn(_worker, args=(world, _free_port(), N, l2, out), nprocs=world, join=True)
    assert np.load(out).shape == (N, N)
