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

    # ---- peer form: shared-memory stand-ins for CUDA IPC buffers, immediate "streams" -------------------------
    class _Shm:
        def __init__(self, shape, dtype, handle=None):
            import mmap
            dt = {"float32": np.float32, "float64": np.float64, "int32": np.int32, "int64": np.int64, "int8": np.int8}[dtype]
            nbytes = max(8, int(np.prod(shape)) * np.dtype(dt).itemsize)
            if handle is None:
                self.fd = os.memfd_create("fuf_test_peer")
                os.ftruncate(self.fd, nbytes)
                self.handle = (os.getpid(), self.fd)
            else:
                self.fd = os.open(f"/proc/{handle[0]}/fd/{handle[1]}", os.O_RDWR)
                self.handle = handle
            self.map = mmap.mmap(self.fd, nbytes)
            self.tensor = torch.from_numpy(np.frombuffer(self.map, dtype=dt, count=int(np.prod(shape))).reshape(shape))

    def peer_alloc(self, shape, dtype):
        return self._Shm(shape, dtype)

    def peer_open(self, handle, shape, dtype):
        return self._Shm(shape, dtype, handle)

    def new_stream(self):
        return None

    def record_event(self, stream=None):
        return None

    def stream_wait(self, stream, event):
        pass

    def zero_(self, t):
        t.zero_()

    def copy_async(self, dst, src, stream):
        self.pulled = getattr(self, "pulled", 0) + 1
        dst.copy_(src)

    def stream_write_u32(self, flags, index, value, stream):
        flags[index] = value

    def profile_matrix(self, rows):
        return torch.zeros((rows, self.n), dtype=torch.float64)

    def host_target(self, D_host):
        return torch.from_numpy(D_host.array)

    def copy_to_target(self, D, target):
        target.copy_(D)

    gram_seg_units = 16

    def stream_wait_u32(self, words, index, value, stream, exact=False):
        if exact:
            # completion counters: the stand-in's "kernels" have run by the time the waits are queued, so the counter must
            # sit exactly at its target
            assert int(words[index]) == value, (index, int(words[index]), value)
        else:
            # band announcements come from another process (shared memory): really wait, bounded
            import time
            t0 = time.time()
            while int(words[index]) - value < 0:
                assert time.time() - t0 < 120, f"band announcement {index} never arrived"
                time.sleep(0.001)
        self.waits = getattr(self, "waits", 0) + 1

    def fill_(self, t, value):
        t.fill_(value)

    def set_sm_limit(self, sms):
        self.sm_limit = sms

    def copy_cols(self, dst, src, c0, c1, stream):
        self.pulled_bands = getattr(self, "pulled_bands", 0) + 1
        dst[:, c0:c1] = src[:, c0:c1]

    def copy_block(self, target, src, rows, cols, stream):
        (r0, r1), (c0, c1) = rows, cols
        assert not bool(torch.isnan(src[r0:r1, c0:c1]).any()), "a forwarded rectangle holds entries no tile has written"
        target[r0:r1, c0:c1] = src[r0:r1, c0:c1]
        src[r0:r1, c0:c1] = float("nan")        # (forwarded once: the next step must produce it again)

    def pairwise_tiles(self, PS, N, mode, tiles, flags, epoch, D, shard, shard_stride, stat, flag_list, list_cap,
                       seg=None, seg_done=None, units=1):
        G, n, sh = PS.shape
        kappa = 1.6e13 if self.l2 else 2e6
        assert (seg is None) == (seg_done is None) and (seg is None or len(seg) == len(tiles))
        for c, (ti, tj, dep) in enumerate(tiles):
            if seg is not None:
                seg_done[seg[c]] += units
            assert dep < 0 or int(flags[dep]) == epoch, "tile scheduled before its shard's flag"
            i0, i1, j0, j1 = ti * TILE, min(N, (ti + 1) * TILE), tj * TILE, min(N, (tj + 1) * TILE)
            A = PS[i0 // sh, :, i0 % sh:i0 % sh + (i1 - i0)].T.double().numpy()
            B = PS[j0 // sh, :, j0 % sh:j0 % sh + (j1 - j0)].T.double().numpy()
            d = np.abs(A[:, None, :] - B[None, :, :])
            blk = ((d * d) if self.l2 else d).sum(-1)
            if ti == tj:
                np.fill_diagonal(blk, 0.0)
            D[i0:i1, j0:j1] = torch.from_numpy(blk)
            D[j0:j1, i0:i1] = torch.from_numpy(blk.T.copy())
            if flag_list is not None:
                st = stat.numpy()
                b = (np.sqrt(st[i0:i1])[:, None] + np.sqrt(st[j0:j1])[None, :]) ** 2 if self.l2 else \
                    st[i0:i1][:, None] + st[j0:j1][None, :]
                gi, gj = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), indexing="ij")
                hit = (blk < kappa * b) & (b != 0) & (gi < gj)
                for a_, b_ in zip(gi[hit], gj[hit]):
                    at = int(flag_list[0])
                    flag_list[0] = at + 1
                    if at < list_cap:
                        flag_list[1 + 2 * at], flag_list[2 + 2 * at] = int(a_), int(b_)

    # --L2 tensor-core path in phases: the "digit planes" of this stand-in are the four bytes of the fp32 operand, so the
    # exchange moves exactly the buffers the CUDA path moves and a tile computed from an un-pulled shard comes out wrong
    def gram_colstats(self, PT_shard, n_valid, stats):
        P = PT_shard[:, :n_valid]
        n = self.n
        stats[:n] = P.abs().max(1).values if n_valid else 0
        stats[n:2 * n] = (P != 0).sum(1).float()
        stats[2 * n:] = (P * P).sum(1)

    def gram_plan(self, stats):
        assert bool((stats[:self.n] >= 0).all())
        self.plan_stats = stats.clone()
        return self.n, 16

    def gram_digits(self, PT_shard, n_valid, planes, shard_index, err_in, q, h2, err_out, SUB):
        raw = PT_shard.contiguous().numpy().view(np.int8).reshape(self.n, PT_shard.shape[1], 4)
        for pl in range(4):
            planes[pl, shard_index, :self.n] = torch.from_numpy(raw[:, :, pl].copy())
        q[:n_valid] = (PT_shard[:, :n_valid].double() ** 2).sum(0)
        h2[:n_valid] = 0
        err_out[:n_valid] = err_in[:n_valid]

    def gram_tiles(self, planes, N, items, flags, epoch, q, h2, err, SUB, D, flag_list, list_cap, seg=None, seg_done=None):
        G, sh = planes.shape[1], planes.shape[3]
        raw = np.ascontiguousarray(planes[:, :, :self.n].permute(1, 2, 3, 0).numpy())            # [G, n, shard, 4] bytes
        PS = torch.from_numpy(raw.view(np.float32).reshape(G, self.n, sh))
        tiles = []
        for c, (ti, b, dep) in enumerate(items):
            if seg is not None:
                seg_done[seg[c]] += self.gram_seg_units
            for tj in (2 * b, 2 * b + 1):
                if tj >= ti and tj * TILE < N:
                    tiles.append((ti, tj, dep))
        l2, self.l2 = self.l2, True
        self.pairwise_tiles(PS, N, 1, tiles, flags, epoch, D, sh, self.n * sh, err, flag_list, list_cap)
        self.l2 = l2

    def refine_list(self, P64S, N, mode, flag_list, list_cap, D, shard, shard_stride):
        self.refine_calls += 1
        P = self._samples(P64S, N, shard)
        for q in range(min(int(flag_list[0]), list_cap)):
            i, j = int(flag_list[1 + 2 * q]), int(flag_list[2 + 2 * q])
            d = np.abs(P[i] - P[j])
            v = float((d * d).sum() if self.l2 else d.sum())
            D[i, j] = v
            D[j, i] = v

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
    def to_device(self, X_host, stage, stream=None):
        src = X_host if torch.is_tensor(X_host) else torch.from_numpy(np.ascontiguousarray(X_host))
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


def _worker(rank, world, port, N, l2, out_path, peer):
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
        if peer and N > 40:                          # a near-duplicate pair across two shards: must be flagged and refined
            X[N - 3] = X[5] * (1 + 1e-9 * rng.random(n))
        X /= X.sum(1, keepdims=True)
        ops = OracleOps(parent, length, l2)
        gram = bool(l2 and N >= 512)
        job = sharded.ShardedAllPairs(ops, N, int(l2), world, rank, dist, peer=peer)
        assert job.use_peer == bool(peer) and (not gram or job.use_gram)
        marks = []
        D = job.step(torch.from_numpy(X[job.lo:job.hi]), marks.append)
        assert marks[0] == "start" and marks[-1] == "end" and "pair_begin" in marks
        if peer:
            # shard pairs: every rank pulled floor(G/2) shards at most, and only those that hold tiles
            # (Gram form: four digit planes + the bypass rows per pulled shard)
            assert getattr(ops, "pulled", 0) == len(job.pulls) * (5 if job.use_gram else 1) and len(job.pulls) <= world // 2
            want_flag = 1 if N > 40 else 0
            assert job.last_flagged == want_flag and ops.refine_calls == want_flag
            assert getattr(ops, "p64_calls", 0) == (want_flag if job.hi > job.lo else 0)
        else:
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
            if peer and N > 40:                         # the refined pair is exact to float64 rounding
                assert abs(got[5, N - 3] - ref[5, N - 3]) <= 1e-12 * ref[5, N - 3] + 1e-18
            assert np.array_equal(got, got.T) and not np.diag(got).any()
            np.save(out_path, got)
        else:
            assert D is None
        # host-buffer variant: every rank stores its rows into one shared host matrix; all ranks see the full result
        shared = sharded.SharedHostMatrix(N, world, rank, dist, register=False)
        shared.array.fill(np.nan)
        dist.barrier()
        Dh = job.step_host(X[job.lo:job.hi], shared)
        if peer and getattr(job, "bp", None) is not None:      # band-wise pulls: every band of every pulled shard, once
            assert getattr(ops, "pulled_bands", 0) == len(job.pulls) * sharded.N_PULL_BANDS
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


@pytest.mark.parametrize("world,N,l2,peer", [(2, 300, False, True), (3, 700, False, True), (4, 400, True, True),
                                             (3, 100, False, True), (2, 300, False, False), (2, 129, True, False),
                                             (3, 900, True, True), (2, 520, True, True), (2, 2100, False, True),
                                             (4, 2300, False, True), (3, 1600, False, True)])
def test_sharded_exchange_gloo(world, N, l2, peer, tmp_path, monkeypatch):
    """peer=True: the shard-pair form (NVLink pulls + in-place stores, here over shared memory); peer=False: the
    collective form (all-gather, tile rows, packed gather) the Gram path still uses."""
    if world == 3:      # once through the named shared-memory file instead of /proc/<pid>/fd
        monkeypatch.setenv("FUF_SHARED_NO_PROCFD", "1")
    if N in (2300, 1600):   # the experimental band-wise pulls of the host-buffer form (off by default)
        monkeypatch.setenv("FUF_BAND_PULLS", "1")
    out = str(tmp_path / "D.npy")
    mp.spawn(_worker, args=(world, _free_port(), N, l2, out, peer), nprocs=world, join=True)
    assert np.load(out).shape == (N, N)


def test_peer_tile_plan_covers_the_triangle_once():
    for N, world in [(10000, 8), (10000, 2), (700, 3), (129, 4), (100, 3), (5000, 5), (128 * 7, 7)]:
        seen, counts = set(), []
        for r in range(world):
            tiles, pulls = sharded.peer_tile_plan(N, world, r)
            counts.append(len(tiles))
            assert len(pulls) <= world // 2 and r not in pulls
            deps = [d for _, _, d in tiles if d >= 0]
            assert sorted(set(deps), key=deps.index) == pulls            # tiles are listed in pull order
            assert all(d < 0 for _, _, d in tiles[:len([t for t in tiles if t[2] < 0])])   # diagonal block first
            sh = sharded.shard_size(N, world) // TILE
            for ti, tj, d in tiles:
                assert ti <= tj and (ti, tj) not in seen
                seen.add((ti, tj))
                assert {ti // sh, tj // sh} <= {r, d if d >= 0 else r}     # operands: own shard and the dependency
        T = (N + TILE - 1) // TILE
        assert len(seen) == T * (T + 1) // 2
        if N == 10000 and world == 8:
            assert max(counts) / (sum(counts) / world) < 1.03


def test_completion_segments_tile_the_matrix_once():
    """The rectangles the completion segments of all ranks forward (sharded.segment_regions) cover every entry of the
    matrix exactly once, for tiles and for the Gram items derived from them; every item inherits one segment."""
    for N, world, mult in [(10000, 2, 128), (10000, 8, 128), (700, 3, 128), (129, 4, 128), (50000, 8, 256), (5000, 5, 256),
                           (2100, 2, 128)]:
        T = (N + TILE - 1) // TILE
        cover = np.zeros((N, N), np.int8) if N <= 10000 else None
        tcover = np.zeros((T, T), np.int32)
        for r in range(world):
            sh = sharded.shard_size(N, world, mult)
            tiles, pulls, seg = sharded.peer_tile_plan(N, world, r, sh, segments=True)
            assert len(seg) == len(tiles) and seg == sorted(seg)            # numbered in processing order
            for regs in sharded.segment_regions(tiles, seg, N):
                for a, b, c, d in regs:
                    assert 0 <= a < b <= N and 0 <= c < d <= N
                    tcover[a // TILE:(b + TILE - 1) // TILE, c // TILE:(d + TILE - 1) // TILE] += 1
                    if cover is not None:
                        cover[a:b, c:d] += 1
            items, item_seg = sharded.gram_items(tiles, seg)
            assert len(items) == len(item_seg) == len(set((i[0], i[1]) for i in items))
        assert (tcover == 1).all() and (cover is None or (cover == 1).all()), (N, world)


def test_banded_pull_plan():
    """Band-wise pulls: every tile of the triangle exactly once over ranks and steps; a tile sits in the step at which both
    of its bands are in, its dependency names the remote band; the completion segments of all ranks tile the matrix once."""
    for N, world, nb in [(10000, 2, 4), (10000, 8, 4), (10000, 4, 4), (2100, 2, 4), (5000, 5, 3), (700, 3, 2), (50000, 8, 4),
                         (1300, 2, 4)]:
        sh = sharded.shard_size(N, world)
        ts, T = sh // TILE, (N + TILE - 1) // TILE
        bands = sharded.shard_bands(sh, nb)
        assert bands[0][0] == 0 and bands[-1][1] == ts and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))

        def band_of(t):
            return [b for b, (lo, hi) in enumerate(bands) if lo <= t % ts < hi][0]

        seen, cover = set(), np.zeros((T, T), np.int32)
        for r in range(world):
            steps, pulls, n_seg = sharded.banded_pull_plan(N, world, r, sh, nb)
            assert pulls == sharded.peer_tile_plan(N, world, r, sh)[1] and len(steps) == nb
            tiles = [t for st in steps for t in st[0]]
            seg = [g for st in steps for g in st[1]]
            assert seg == sorted(seg) and (not seg or seg[-1] == n_seg - 1)
            for s, (tl, _) in enumerate(steps):
                for ti, tj, dep in tl:
                    assert ti <= tj and (ti, tj) not in seen
                    seen.add((ti, tj))
                    if dep < 0:
                        assert ti // ts == r == tj // ts and band_of(tj) == s
                    else:
                        q, br = divmod(dep, nb)
                        own_t, rem_t = (ti, tj) if ti // ts == r else (tj, ti)
                        assert q in pulls and rem_t // ts == q and own_t // ts == r
                        assert band_of(rem_t) == br and max(band_of(own_t), br) == s
            for regs in sharded.segment_regions(tiles, seg, N):
                for a, b, c, d in regs:
                    cover[a // TILE:(b + TILE - 1) // TILE, c // TILE:(d + TILE - 1) // TILE] += 1
        assert len(seen) == T * (T + 1) // 2 and (cover == 1).all(), (N, world)
