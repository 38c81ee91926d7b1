"""Multi-GPU all-pairs: one process per GPU of one node, samples sharded, the distance matrix written in place.

The distance matrix has no cross-pair dependence (every ``dists[i, j]`` of the reference loop,
``src/algorithms/emd_unifrac.py:49-54``, needs only rows i and j of the pushed profiles), so the job shards
without any collective inside the arithmetic.  Weighted L1 (and the direct --L2 kernel):

1. rank r pushes up its own block of ``shard`` samples              (``fuf_push_up`` -> ``[n, shard]`` node-major, in a
   buffer its peers have mapped through CUDA IPC)
2. the per-sample statistics are all-gathered (NCCL, 80 KB)          — which is also the signal that every shard is ready
3. the tile grid is dealt by SHARD PAIRS (``peer_tile_plan``): rank r owns its diagonal block and the blocks
   (r, r+1) .. (r, r+G/2) (the last one split with its partner), so it needs only G/2 of the G-1 remote shards.  A copy
   stream pulls them from the peers' buffers over NVLink, one after the other, and raises a flag behind each copy.
4. ONE launch of ``fuf_pairwise_tiles`` walks the rank's tiles in arrival order — diagonal block first, which needs
   nothing remote; a tile whose shard is still in flight makes the kernel's producer warp wait for the flag.  Rank 0
   stores every tile, and its mirror, straight into its own matrix.  Every other rank — and every rank when the result
   goes to host memory (``step_host``) — stores into a matrix of its own, and a copy engine forwards each finished
   *completion segment* (one rasterisation strip of a shard-pair block: its rectangle and the mirrored one,
   ``segment_regions``) into rank 0's matrix over NVLink, or into the page-locked host matrix the ranks share over the
   rank's own PCIe link, while the launch works on the next segment: the kernel counts finished tiles per segment, the
   copy stream waits on the counter (``fuf_stream_wait_u32``) and issues pitched copies (``fuf_copy2d_async``).
   Nothing is packed, gathered or assembled, and no kernel stores across NVLink or PCIe.
   The refinement criterion is evaluated in the tile epilogues; the flagged-pair count is all-reduced (NCCL, 8 bytes:
   the end-of-step barrier), and only if it is non-zero are the float64 profiles pushed up, exchanged and the listed
   pairs recomputed (``fuf_refine_list``).

--L2 from 512 samples on (tensor-core Gram kernel) takes the same form in phases: column statistics of the own shard,
all-reduced, so that every rank derives the same plan (``fuf_gram_colstats`` / ``fuf_gram_plan``); int8 digit planes of
the own shard only (``fuf_gram_digits``); the copy stream pulls the planes — not the fp32 operands — of the G/2 shards
the rank's tiles touch; one launch of the CTA-pair tile kernel (``fuf_gram_tiles``) with the same arrival flags,
completion segments and epilogue flags.  ``peer=False`` keeps the older collective form (all-gather of the pushed blocks, tile rows
dealt boustrophedon-wise, packed gather to rank 0 or per-rank stores into the shared host matrix).

The compute calls are injected (``ops``) so that the partition / exchange logic is testable on CPU with the
``gloo`` backend; the product always injects ``engine.TreeContext`` (CUDA).  ``world == 1`` needs no process group.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence  # noqa: F401

TILE = 128
# tile rows per rasterisation strip: floor(sqrt(148 SMs)); FUF_STRIP overrides it for experiments (1000 = row-major blocks)
STRIP = int(os.environ.get("FUF_STRIP", "12"))
N_PULL_BANDS = 4        # bands per shard of the host-buffer form with band-wise pulls


def shard_size(N: int, world: int, multiple: int = TILE) -> int:
    """Samples per rank: ceil(N / world) rounded up to a multiple of the 128-sample tile (256 for the CTA-pair Gram kernel
    of the peer form, whose work items are pairs of column tiles)."""
    per = (N + world - 1) // world
    return max(multiple, (per + multiple - 1) // multiple * multiple)


def sample_range(N: int, world: int, rank: int, shard: Optional[int] = None):
    s = shard or shard_size(N, world)
    return min(N, rank * s), min(N, (rank + 1) * s)


def gram_items(tiles, seg=None):
    """(ti, tj, dependency) tiles of ``peer_tile_plan`` -> work items (ti, b, dependency) of the CTA-pair Gram kernel: row
    tile ti against the column-tile pair (2b, 2b + 1), in the same order (shards are multiples of 256, so a pair never
    straddles two shards).  With ``seg`` (the completion segment of every tile) also the segment of every item."""
    items, item_seg, seen = [], [], set()
    for c, (ti, tj, dep) in enumerate(tiles):
        key = (ti, tj // 2)
        if key not in seen:
            seen.add(key)
            items.append((ti, tj // 2, dep))
            if seg is not None:
                item_seg.append(seg[c])
    return items if seg is None else (items, item_seg)


def segment_regions(tiles, seg, N: int):
    """Rectangles of the full matrix that a completion segment finishes, per segment: [(row0, row1, col0, col1), ...] in
    samples.  A segment's tiles fill the upper-triangle part of their bounding box of tiles (checked); the kernel stores
    every tile and its mirror, so the finished entries are the box itself and its transposed image — cut so that the
    rectangles of one segment do not overlap (strips and column bands of a diagonal block meet the diagonal)."""
    groups = {}
    for (ti, tj, _), sg in zip(tiles, seg):
        groups.setdefault(sg, []).append((ti, tj))
    assert sorted(groups) == list(range(len(groups))), "segments are numbered 0 .. S-1"

    def rect(a, b, c, d):
        return (a * TILE, min(N, b * TILE), c * TILE, min(N, d * TILE))

    out = []
    for sg in range(len(groups)):
        g = groups[sg]
        ra, rb = min(t[0] for t in g), max(t[0] for t in g) + 1
        ca, cb = min(t[1] for t in g), max(t[1] for t in g) + 1
        want = sum(1 for ti in range(ra, rb) for tj in range(max(ca, ti), cb))
        assert len(set(g)) == len(g) == want, f"segment {sg} does not fill its box of tiles"
        regs = [rect(ra, rb, ca, cb),
                rect(max(ca, ra), min(cb, rb), ra, min(rb, ca)),        # mirror rows inside the box's rows: left of it
                rect(max(ca, rb), cb, ra, rb)]                          # mirror rows below the box
        out.append([r for r in regs if r[1] > r[0] and r[3] > r[2]])
    return out


def tile_row_owner(t: int, world: int) -> int:
    """Boustrophedon deal of tile rows: row t of the upper triangle holds T - t tiles, so dealing
    0,1,..,G-1,G-1,..,1,0,0,1,.. keeps every rank within one tile row of the mean tile count."""
    q = t % (2 * world)
    return q if q < world else 2 * world - 1 - q


def owned_tile_rows(N: int, world: int, rank: int) -> List[int]:
    T = (N + TILE - 1) // TILE
    return [t for t in range(T) if tile_row_owner(t, world) == rank]


def packed_offsets(rows: Sequence[int], N: int):
    """Packed layout of a rank's row blocks: slot c keeps [128, N - 128 t] (the columns its tile row owns), contiguous,
    at offsets[c] doubles.  Returns (offsets, total); -1 rows are padding slots and take no room."""
    offsets, total = [], 0
    for t in rows:
        offsets.append(total)
        if t >= 0:
            total += TILE * (N - TILE * t)
    return offsets, total


def peer_tile_plan(N: int, world: int, rank: int, shard: Optional[int] = None, segments: bool = False):
    """Tiles of rank ``rank`` when the upper-triangular tile grid is dealt by shard pairs, in processing order, as
    (ti, tj, dependency) triples, plus the remote shards the rank pulls, in pull order.

    Shard A holds tile indices [A Ts, (A+1) Ts).  Rank r owns the diagonal block (r, r) — no remote operand — then the
    blocks of shard pair {r, (r + d) mod G} for d = 1 .. G/2: a round-robin tournament, every unordered pair has exactly
    one owner, except that for even G the pairs at distance G/2 are split by tile rows between their two members.  Every
    rank gets the same number of tiles (up to the ragged last shard) and pulls only floor(G/2) of the G - 1 remote shards.
    ``segments=True`` adds a third value: the completion segment of every tile — one per rasterisation strip, numbered in
    processing order (``segment_regions`` gives the part of the matrix each one finishes).
    """
    shard = shard or shard_size(N, world)
    Ts, T = shard // TILE, (N + TILE - 1) // TILE

    def tiles_of(A):
        return range(A * Ts, min((A + 1) * Ts, T))

    seg = []

    def strips(rows, cols, dep):
        # L2 rasterisation (see pairwise_f32): strips of STRIP tile rows walked column by column, so that the ~148 tiles
        # in flight at any time form a compact block sharing ~2 STRIP operand panels
        out = []
        for r0 in range(0, len(rows), STRIP):
            band = rows[r0:r0 + STRIP]
            part = [(ti, tj, dep) for tj in cols for ti in band if ti <= tj]
            if part:
                seg.extend([seg[-1] + 1 if seg else 0] * len(part))
                out += part
        return out

    own = list(tiles_of(rank))
    tiles = strips(own, own, -1)
    pulls = []
    for d in range(1, world // 2 + 1):
        q = (rank + d) % world
        if q == rank:
            continue
        A, B = min(rank, q), max(rank, q)
        rows = list(tiles_of(A))
        if 2 * d == world:                       # the pair at distance G/2 is shared: lower rank takes the first half
            half = (len(rows) + 1) // 2
            rows = rows[:half] if rank == A else rows[half:]
        block = strips(rows, list(tiles_of(B)), q)
        if block:
            pulls.append(q)
            tiles += block
    return (tiles, pulls, seg) if segments else (tiles, pulls)


def shard_bands(shard: int, n_bands: int):
    """Bands of whole tiles of one shard, the same for every rank: [(first tile, end tile)] relative to the shard's first
    tile.  (The last shard may be ragged: its later bands then hold fewer valid tiles, or none.)"""
    ts = shard // TILE
    return [(ts * b // n_bands, ts * (b + 1) // n_bands) for b in range(n_bands)]


def banded_pull_plan(N: int, world: int, rank: int, shard: int, n_bands: int):
    """The rank's tiles of ``peer_tile_plan`` regrouped for a job whose shards arrive band by band (every rank uploads and
    pushes up its own shard in ``n_bands`` bands and announces each band to the ranks that need it).

    A tile needs one band of the own shard and — off the diagonal block — one band of a remote shard; it joins step
    max(own band, remote band), the first moment both are in.  Returns (steps, pulls, n_segments): ``steps[s]`` is the launch
    of step s as (tiles, seg) with tiles = (ti, tj, dependency) and dependency = q * n_bands + remote band (or -1: the
    diagonal block); ``seg`` numbers the completion segments in launch order, one per box of tiles: the column band of the
    diagonal block, and per remote shard the boxes {own band s} x {remote bands <= s} and {own bands < s} x {remote band s}."""
    ts = shard // TILE
    T = (N + TILE - 1) // TILE
    bands = shard_bands(shard, n_bands)

    def band_of(t):                      # band of global tile index t inside its shard
        local = t % ts
        for b, (lo, hi) in enumerate(bands):
            if lo <= local < hi:
                return b
        raise AssertionError(t)

    tiles, pulls = peer_tile_plan(N, world, rank, shard)
    steps = [([], []) for _ in range(n_bands)]
    boxes = {}                           # (step, remote shard or -1, which box) -> tiles
    for ti, tj, dep in tiles:
        if dep < 0:
            boxes.setdefault((band_of(tj), -1, 0), []).append((ti, tj, -1))
            continue
        own_t, rem_t = (ti, tj) if ti // ts == rank else (tj, ti)
        bo, br = band_of(own_t), band_of(rem_t)
        step = max(bo, br)
        boxes.setdefault((step, dep, 0 if bo == step else 1), []).append((ti, tj, dep * n_bands + br))
    n_seg = 0
    order = [-1] + pulls
    for s in range(n_bands):
        for q in order:
            for which in (0, 1):
                box = boxes.get((s, q, which))
                if not box:
                    continue
                # rasterise the box like peer_tile_plan does: strips of STRIP tile rows, column by column
                rows = sorted(set(t[0] for t in box))
                by = {}
                for t in box:
                    by.setdefault((t[0], t[1]), t)
                cols = sorted(set(t[1] for t in box))
                for r0 in range(0, len(rows), STRIP):
                    strip = rows[r0:r0 + STRIP]
                    part = [by[(ti, tj)] for tj in cols for ti in strip if (ti, tj) in by]
                    steps[s][0].extend(part)
                    steps[s][1].extend([n_seg] * len(part))
                n_seg += 1
    assert sum(len(st[0]) for st in steps) == len(tiles) and T >= 0
    return steps, pulls, n_seg


def refine_budget(N: int) -> int:
    """Pairs worth recomputing one by one (0.1 % of all pairs, at least 1,024); more than that and the float64 tile kernel
    is cheaper."""
    return max(1024, N * (N - 1) // 2 // 1000)


def tiles_of(rows: Sequence[int], N: int) -> int:
    T = (N + TILE - 1) // TILE
    return sum(T - t for t in rows)


class ShardedAllPairs:
    """One rank's part of an all-pairs job over ``N`` samples.

    ``ops`` provides ``n``, ``empty(shape, dtype)``, ``push_up_center``, ``push_up``, ``pairwise``, ``refine``,
    ``pack_tile_rows``, ``assemble_packed`` (and the host-buffer calls) with the semantics of the C ABI; ``dist`` is ``torch.distributed`` (None when world == 1).
    ``refine=True`` also exchanges the float64 profiles and runs the float64 refinement pass on this rank's rows.
    """

    def __init__(self, ops, N: int, mode: int, world: int = 1, rank: int = 0, dist=None, refine: bool = True,
                 gram: Optional[bool] = None, peer: Optional[bool] = None):
        self.ops, self.N, self.mode, self.world, self.rank, self.dist = ops, N, mode, world, rank, dist
        self.do_refine = refine
        # --L2 from 512 samples on: tensor-core Gram kernel (its refinement needs the float64 profiles)
        self.use_gram = (mode == 1 and N >= 512 and refine) if gram is None else bool(gram)
        # peer form on several GPUs (shard pairs, NVLink pulls, in-place stores): weighted L1, the direct --L2 kernel and the
        # Gram kernel; peer=False keeps the collective form (all-gather, tile rows, packed gather)
        self.use_peer = world > 1 and peer is not False
        self.shard = shard_size(N, world, 2 * TILE if (self.use_peer and self.use_gram) else TILE)
        self.lo, self.hi = sample_range(N, world, rank, self.shard)
        self.all_rows = [owned_tile_rows(N, world, r) for r in range(world)]
        self.rows = self.all_rows[rank]
        self.max_rows = max(len(r) for r in self.all_rows)
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs a torch.distributed process group")
        n = ops.n
        self.epoch = 0
        if self.use_peer:
            self._init_peer()
            return
        self.PS = ops.empty((world, n, self.shard), "float32")          # pushed blocks of all ranks (zero-padded)
        self.P64S = None        # float64 pushed blocks: allocated, computed and exchanged only if a pair is flagged
        self.stat = ops.empty((world * self.shard,), "float64")
        self.center = ops.empty((n,), "float64")
        # full matrix on GPU 0: needed by step(); step_host() with several ranks never touches it (allocated lazily)
        self.D = ops.empty((N, N), "float64") if world == 1 else None
        self.blocks = self.gathered = None
        self.last_flagged = 0
        self._x_stage = None
        self._pending = None
        self._x_local = None
        self.flagged = ops.empty((1,), "int64")
        if world > 1:
            self.blocks = ops.empty((self.max_rows * TILE, N), "float64")
            # what crosses NVLink to rank 0: each rank's row blocks packed to the columns they own (half the volume)
            self.offs, _ = packed_offsets(self.rows, N)
            self.max_packed = max(1, max(packed_offsets(r, N)[1] for r in self.all_rows))
            self.packed = None
            self.slot_rows, self.slot_offs = [], []
            for r in range(world):
                self.slot_rows += self.all_rows[r]
                self.slot_offs += [r * self.max_packed + o for o in packed_offsets(self.all_rows[r], N)[0]]

    def pushup_bytes_per_entry(self, x_itemsize: int) -> int:
        """HBM bytes the push-up of this job has to move per (sample, node): the profile entry in, the centred fp32
        entry out (the float64 profiles of the refinement pass are only produced when a pair is flagged)."""
        return x_itemsize + 4

    def release_device_matrix(self):
        """Drop the device-resident result matrix of ``step`` (world == 1); ``step`` allocates it again on demand."""
        self.D = None

    def _pairwise(self, rows, D):
        """Distances of this rank's tile rows (all rows when ``rows`` is None) into D; returns the refinement
        statistics of the refinement scan to run afterwards: (rounding statistic, linear statistic or None)."""
        ops, N, sh, stride = self.ops, self.N, self.shard, self.ops.n * self.shard
        if self.use_gram:
            q, err = ops.pairwise_gram(self.PS, N, self.stat, rows, sh, stride, D)
            return err, q                        # rounding criterion + linear criterion on the Gram norms
        ops.pairwise(self.PS, N, self.mode, rows, sh, stride, D)
        return self.stat, None

    def _compute(self, X_local, mark):
        """Phases 1-3 (+ refinement).  Afterwards ``self.D`` (world == 1) or ``self.blocks`` (this rank's tile rows, upper
        part) hold the final distances."""
        ops, N, mode, r, sh, n = self.ops, self.N, self.mode, self.rank, self.shard, self.ops.n
        stride = n * sh
        mark("start")
        if r == 0:
            ops.push_up_center(X_local, mode, self.center)
        if self.world > 1:
            self.dist.broadcast(self.center, src=0)
        mark("push_begin")      # (after the centre: its kernels and the broadcast)
        self._x_local = X_local
        if self.hi > self.lo:
            ops.push_up(X_local, mode, self.PS[r], 0, self.center, None, self.stat[r * sh:(r + 1) * sh])
        mark("push_end")
        self._pending = None
        if self.world == 1:
            if self.D is None:
                self.D = ops.empty((N, N), "float64")
            mark("pair_begin")
            stat, lin = self._pairwise(None, self.D)
            mark("pair_end")
            if self.do_refine and N > 1:
                ops.count_uncertified_into(N, mode, stat, self.D, None, lin, self.flagged)
                self._pending = (stat, lin)
            mark("refine_end")
            return
        self.dist.all_gather_into_tensor(self.PS.view(-1), self.PS[r].reshape(-1))
        self.dist.all_gather_into_tensor(self.stat, self.stat[r * sh:(r + 1) * sh])
        mark("pair_begin")
        stat, lin = self._pairwise(self.rows, self.blocks) if self.rows else (None, None)
        mark("pair_end")
        if self.do_refine:
            # The float64 profiles (2x the fp32 volume) are pushed up and cross NVLink only if some rank has a pair to
            # refine.  The count stays on the device: it is all-reduced in stream order and the caller goes on queueing
            # the collection of the results; the host looks at the count afterwards (_refine_if_flagged) and, in the
            # rare case, refines and collects again.
            if self.rows:
                ops.count_uncertified_into(N, mode, stat, self.blocks, self.rows, lin, self.flagged)
            else:
                self.flagged.zero_()
            self.dist.all_reduce(self.flagged)
            self._pending = (stat, lin)
        mark("refine_end")

    def _refine_if_flagged(self) -> bool:
        """Second half of the refinement decision (one host synchronisation).  True: rows were refined, collect again."""
        self.last_flagged = 0
        if self._pending is None:
            return False
        stat, lin = self._pending
        self._pending = None
        self.last_flagged = int(self.flagged.item())
        if self.last_flagged == 0:
            return False
        ops, r, sh, n = self.ops, self.rank, self.shard, self.ops.n
        if self.P64S is None:
            self.P64S = ops.empty((self.world, n, sh), "float64")
        if self.hi > self.lo:
            ops.push_up_p64(self._x_local, self.mode, self.P64S[r])
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.P64S.view(-1), self.P64S[r].reshape(-1))
        rows, target = (self.rows, self.blocks) if self.world > 1 else (None, self.D)
        if self.world == 1 or self.rows:
            # per-pair recomputation costs ~1 us, the float64 tile kernel ~3 ns per pair: beyond 0.1 % of the pairs
            # (the count is global, every rank takes the same branch) the float64 kernel redoes this rank's rows —
            # the same budget rule as the one-GPU paths (engine._refine_or_redo, fuf_all_pairs_host)
            if self.last_flagged > refine_budget(self.N):
                ops.pairwise_f64(self.P64S, self.N, self.mode, rows, sh, n * sh, target)
            else:
                ops.refine(self.P64S, self.N, self.mode, stat, target, rows, sh, n * sh, lin)
        return True

    # ------------------------------------------------------------------------------------------------------
    # peer form
    # ------------------------------------------------------------------------------------------------------
    def _init_peer(self):
        ops, N, world, rank, dist, sh, n = self.ops, self.N, self.world, self.rank, self.dist, self.shard, self.ops.n
        self.tiles, self.pulls, self.seg = peer_tile_plan(N, world, rank, sh, segments=True)
        self.rows = self.tiles                    # (truthy iff this rank computes anything)
        # every rank's operand buffer [G, n, shard]: slot r is written by its own push-up and read by the peers' copy
        # engines; the other slots receive the pulled shards
        self.D_buf = ops.peer_alloc((N, N), "float64") if rank == 0 else None
        handles = [None] * world
        if self.use_gram:
            # --L2 on the tensor cores: what crosses NVLink are the int8 digit planes (and the few float64 bypass rows)
            # of the shards this rank's tiles touch; the fp32 operands never leave their GPU
            self.items, self.item_seg = gram_items(self.tiles, self.seg)
            self.rows_cap = (n + 4096 + 63) // 64 * 64            # plan rows: n + padding of the scale buckets
            self.sub_cap = 256                                    # float64 bypass rows (GRAM_DIRECT_ROWS)
            self.PT_own = ops.empty((n, sh), "float32")
            self.planes_buf = ops.peer_alloc((4, world, self.rows_cap, sh), "int8")
            self.sub_buf = ops.peer_alloc((world, self.sub_cap, sh), "float64")
            self.planes, self.SUB = self.planes_buf.tensor, self.sub_buf.tensor
            self.gq, self.gh2, self.gerr = (ops.empty((world * sh,), "float64") for _ in range(3))
            self.colstats = ops.empty((3 * n,), "float32")
            dist.all_gather_object(handles, (self.planes_buf.handle, self.sub_buf.handle,
                                             self.D_buf.handle if rank == 0 else None))
            self.peer_planes = {q: ops.peer_open(handles[q][0], (4, world, self.rows_cap, sh), "int8") for q in self.pulls}
            self.peer_sub = {q: ops.peer_open(handles[q][1], (world, self.sub_cap, sh), "float64") for q in self.pulls}
        else:
            self.PS_buf = ops.peer_alloc((world, n, sh), "float32")
            self.PS = self.PS_buf.tensor
            # experimental band-wise pulls of the host-buffer form (FUF_BAND_PULLS=1): the statistics travel with the bands
            # (peer-readable), and every rank announces a pushed-up band by a 4-byte copy into the `ready` words of the
            # ranks that pull its shard
            self.band_pulls = os.environ.get("FUF_BAND_PULLS", "0") == "1" and sh // TILE >= N_PULL_BANDS
            self.stat_buf = ops.peer_alloc((world * sh,), "float64") if self.band_pulls else None
            self.ready_buf = ops.peer_alloc((world * N_PULL_BANDS,), "int32") if self.band_pulls else None
            dist.all_gather_object(handles, (self.PS_buf.handle, self.stat_buf.handle if self.band_pulls else None,
                                             self.ready_buf.handle if self.band_pulls else None,
                                             self.D_buf.handle if rank == 0 else None))
            self.peer_PS = {q: ops.peer_open(handles[q][0], (world, n, sh), "float32") for q in self.pulls}
            if self.band_pulls:
                self.peer_stat = {q: ops.peer_open(handles[q][1], (world * sh,), "float64") for q in self.pulls}
                self.consumers = [q for q in range(world) if q != rank and rank in peer_tile_plan(N, world, q, sh)[1]]
                self.peer_ready = {q: ops.peer_open(handles[q][2], (world * N_PULL_BANDS,), "int32") for q in self.consumers}
        if rank != 0:
            self.D_buf = ops.peer_open(handles[0][-1], (N, N), "float64")
        self.D = self.D_buf.tensor if rank == 0 else None
        self.D_target = self.D_buf.tensor         # rank 0's matrix, as seen from this rank
        self.P64S = None
        self.stat = self.stat_buf.tensor if (not self.use_gram and self.band_pulls) else ops.empty((world * sh,), "float64")
        self.center = ops.empty((n,), "float64")
        self.flags = ops.empty((world,), "int32")                         # arrival flag per shard
        self.list_cap = refine_budget(N)
        self.flag_list = ops.empty((1 + 2 * self.list_cap,), "int64")     # [0] = count, then (i, j) pairs
        self.flagged = ops.empty((1,), "int64")
        self.copy_stream = ops.new_stream()
        self.d2h_stream = ops.new_stream()
        self.last_flagged = 0
        self._x_stage = self._pending = self._x_local = None
        # A rank whose result matrix lives elsewhere (rank 0's GPU for ranks > 0, the shared host matrix for everybody)
        # computes into a matrix of its own, D_local, and a copy engine moves every finished completion segment to its
        # place while the tile kernel works on the next one: the kernel never stores across NVLink or PCIe.
        self.D_local = None
        units = self.ops.gram_seg_units if self.use_gram else 1
        work, work_seg = (self.items, self.item_seg) if self.use_gram else (self.tiles, self.seg)
        self.plan = _SegmentPlan(work, work_seg, segment_regions(self.tiles, self.seg, N), units)
        # host-buffer form (weighted L1): the own shard is uploaded in up to 4 bands of whole tiles; band b carries the
        # diagonal-block tiles whose column tile lies in it (their rows lie in bands <= b) as ONE segment, so the
        # finished part of the diagonal block leaves for the host while the next band is still coming in
        own_tiles = [t for t in self.tiles if t[2] < 0]
        self.offdiag_tiles = [t for t in self.tiles if t[2] >= 0]
        Ts_own = (self.hi - self.lo + TILE - 1) // TILE
        self.diag_bands, self.plan_banded = [], None
        if Ts_own >= 8 and not self.use_gram:
            t0 = rank * (sh // TILE)
            tiles_b, seg_b = [], []
            for b in range(4):
                lo_t, hi_t = t0 + Ts_own * b // 4, t0 + Ts_own * (b + 1) // 4
                s0, s1 = (lo_t - t0) * TILE, min(self.hi - self.lo, (hi_t - t0) * TILE)
                band = [t for t in own_tiles if lo_t <= t[1] < hi_t]
                self.diag_bands.append((s0, s1, len(tiles_b), len(tiles_b) + len(band)))
                tiles_b += band
                seg_b += [b] * len(band)
            first_off = min([sg for t, sg in zip(self.tiles, self.seg) if t[2] >= 0], default=0)
            self.offdiag_from = len(tiles_b)
            tiles_b += self.offdiag_tiles
            seg_b += [4 + sg - first_off for t, sg in zip(self.tiles, self.seg) if t[2] >= 0]
            self.plan_banded = _SegmentPlan(tiles_b, seg_b, segment_regions(tiles_b, seg_b, N), 1)
        # host-buffer form with band-wise pulls (weighted L1, direct --L2): see _compute_peer_band_pulls.  EXPERIMENTAL, off
        # unless FUF_BAND_PULLS=1: correct (gloo emulation with world 2/3/4, 2-3 ranks sharing one B200, two real B200s) but
        # not yet faster — its strided band pulls run as SM kernels on the few SMs left free for them (par. 8).
        self.bp = None
        nb = N_PULL_BANDS
        if not self.use_gram and self.band_pulls:
            steps, _, _ = banded_pull_plan(N, world, rank, sh, nb)
            work = [t for st in steps for t in st[0]]
            work_seg = [g for st in steps for g in st[1]]
            plan = _SegmentPlan(work, work_seg, segment_regions(work, work_seg, N), 1)
            step_range, seg_range, at = [], [], 0
            for tl, sg in steps:
                step_range.append((at, at + len(tl)))
                seg_range.append((sg[0], sg[-1] + 1) if sg else (0, 0))
                at += len(tl)
            cols = [(lo * TILE, hi * TILE) for lo, hi in shard_bands(sh, nb)]            # columns of a shard buffer
            rows_own = self.hi - self.lo
            self.bp = {"plan": plan, "step_range": step_range, "seg_range": seg_range, "cols": cols,
                       "own": [(min(rows_own, c0), min(rows_own, c1)) for c0, c1 in cols]}    # rows of the own samples
            self.ready = self.ready_buf.tensor
            self.epoch_word = ops.empty((1,), "int32")
            self.band_flags = ops.empty((world * nb,), "int32")
        self.seg_done = ops.empty((max(1, self.plan.n_seg, self.plan_banded.n_seg if self.plan_banded else 0,
                                       self.bp["plan"].n_seg if self.bp else 0),), "int32")
        dist.barrier()

    def _local_matrix(self, target):
        """Where the tile kernels store: the target itself when it is this device's memory (rank 0 of ``step``), else the
        rank's own matrix, from which the copy engine forwards the finished segments."""
        if target is self.D or os.environ.get("FUF_NO_FORWARD"):      # (debug switch: the kernels store in place everywhere)
            return target
        if self.D_local is None:
            self.D_local = self.ops.empty((self.N, self.N), "float64")
        return self.D_local

    def _forward_segments(self, plan, D_comp, target, first_seg=0, last_seg=None):
        """Queue, on the device-to-host stream, for every completion segment: wait for its counter, then one pitched copy
        per finished rectangle from the rank's matrix to the target (peer GPU or page-locked host memory)."""
        ops = self.ops
        for sg in range(first_seg, plan.n_seg if last_seg is None else last_seg):
            ops.stream_wait_u32(self.seg_done, sg, plan.expected[sg], self.d2h_stream, exact=True)
            for r0, r1, c0, c1 in plan.regions[sg]:
                ops.copy_block(target, D_comp, (r0, r1), (c0, c1), self.d2h_stream)

    def _begin_segments(self):
        """Counters to zero in compute-stream order; the copy stream may look at them from here on."""
        self.ops.zero_(self.seg_done)
        self.ops.stream_wait(self.d2h_stream, self.ops.record_event())

    def _end_segments(self):
        """The compute stream (and with it the end-of-step all-reduce) waits for the last forwarded segment."""
        self.ops.stream_wait(None, self.ops.record_event(self.d2h_stream))

    def _compute_peer_gram(self, X_local, D_target, mark):
        """The peer form of the --L2 tensor-core path: push-up and column statistics of the own shard, all-reduce of the
        statistics (every rank derives the same plan), digits of the own shard, pulls of the peers' digit planes, one
        launch of the CTA-pair tile kernel storing in place."""
        ops, N, mode, r, sh, n, dist = self.ops, self.N, self.mode, self.rank, self.shard, self.ops.n, self.dist
        self.epoch += 1
        mark("start")
        if r == 0:
            ops.push_up_center(X_local, mode, self.center)
        dist.broadcast(self.center, src=0)
        mark("push_begin")
        self._x_local = X_local
        own = slice(r * sh, (r + 1) * sh)
        n_valid = self.hi - self.lo
        if n_valid:
            ops.push_up(X_local, mode, self.PT_own, 0, self.center, None, self.stat[own])
        mark("push_end")
        ops.gram_colstats(self.PT_own, n_valid, self.colstats)
        dist.all_reduce(self.colstats[:n], op=dist.ReduceOp.MAX)
        dist.all_reduce(self.colstats[n:])
        n_rows, n_sub = ops.gram_plan(self.colstats)              # (one host synchronisation: the plan is data-dependent)
        if n_rows > self.rows_cap or n_sub > self.sub_cap:
            raise RuntimeError(f"Gram plan needs {n_rows} rows / {n_sub} bypass rows, buffers hold {self.rows_cap} / {self.sub_cap}")
        ops.gram_digits(self.PT_own, n_valid, self.planes, r, self.stat[own], self.gq[own], self.gh2[own], self.gerr[own],
                        self.SUB)
        # per-sample norms and statistics of all samples; the last all-gather also says: every shard's planes are ready
        dist.all_gather_into_tensor(self.gq, self.gq[own])
        dist.all_gather_into_tensor(self.gh2, self.gh2[own])
        dist.all_gather_into_tensor(self.gerr, self.gerr[own])
        ops.zero_(self.flag_list[:1])
        D_comp = self._local_matrix(D_target)
        forward = D_comp is not D_target
        if forward:
            self._begin_segments()
        ready = ops.record_event()
        ops.stream_wait(self.copy_stream, ready)
        for q in self.pulls:
            for pl in range(4):
                ops.copy_async(self.planes[pl, q, :n_rows], self.peer_planes[q].tensor[pl, q, :n_rows], self.copy_stream)
            ops.copy_async(self.SUB[q, :n_sub], self.peer_sub[q].tensor[q, :n_sub], self.copy_stream)
            ops.stream_write_u32(self.flags, q, self.epoch, self.copy_stream)
        mark("pair_begin")
        if self.items:
            ops.gram_tiles(self.planes, N, self.items, self.flags, self.epoch, self.gq, self.gh2, self.gerr, self.SUB, D_comp,
                           self.flag_list if self.do_refine else None, self.list_cap,
                           self.plan.seg if forward else None, self.seg_done if forward else None)
            if forward:
                self._forward_segments(self.plan, D_comp, D_target)
                self._end_segments()
        mark("pair_end")
        self.flagged.copy_(self.flag_list[:1])
        dist.all_reduce(self.flagged)
        self._pending = True
        mark("refine_end")

    def _compute_peer(self, X_local, D_target, mark, host_src=None):
        """Steps 1-4 of the module docstring; afterwards the matrix behind ``D_target`` holds this rank's tiles (the
        refinement decision is pending in ``self.flagged``).

        ``host_src``: this rank's un-pushed samples in page-locked host memory (``step_host``).  They are then uploaded
        into ``X_local`` in a few bands of whole tiles; each band is pushed up as soon as it has landed and followed by the
        tiles of the rank's DIAGONAL block whose columns it completes — the block that needs no remote shard is computed
        under the rank's own host-to-device transfer, and the peers' shards are pulled when everybody's upload is done."""
        if self.use_gram:
            if host_src is not None:
                self.ops.to_device(host_src, X_local)
            return self._compute_peer_gram(X_local, D_target, mark)
        if host_src is not None and self.bp is not None:
            return self._compute_peer_band_pulls(X_local, D_target, mark, host_src)
        ops, N, mode, r, sh, n, dist = self.ops, self.N, self.mode, self.rank, self.shard, self.ops.n, self.dist
        self.epoch += 1
        own = slice(r * sh, (r + 1) * sh)
        stat = self.stat if self.do_refine else None
        flist = self.flag_list if self.do_refine else None
        rows_own = self.hi - self.lo
        ops.zero_(self.flag_list[:1])
        D_comp = self._local_matrix(D_target)
        forward = D_comp is not D_target
        banded = not (host_src is None or rows_own == 0 or not self.diag_bands)
        plan = self.plan_banded if banded else self.plan
        seg_done = self.seg_done if forward else None
        if forward:
            self._begin_segments()
        mark("start")
        if not banded:
            if host_src is not None:
                ops.to_device(host_src, X_local)
            if r == 0:
                ops.push_up_center(X_local, mode, self.center)
            dist.broadcast(self.center, src=0)
            mark("push_begin")
            if rows_own:
                ops.push_up(X_local, mode, self.PS[r], 0, self.center, None, self.stat[own])
            mark("push_end")
            rest, rest_seg, seg_from = plan.work, plan.seg, 0
        else:
            events = []
            for s0, s1, _, _ in self.diag_bands:        # all uploads are queued up front on the copy stream
                ops.to_device(host_src[s0:s1], X_local[s0:s1], self.copy_stream)
                events.append(ops.record_event(self.copy_stream))
            ops.stream_wait(None, events[0])
            if r == 0:
                ops.push_up_center(X_local[:self.diag_bands[0][1]], mode, self.center)
            dist.broadcast(self.center, src=0)
            mark("push_begin")
            for b, ((s0, s1, t0, t1), ev) in enumerate(zip(self.diag_bands, events)):
                ops.stream_wait(None, ev)
                ops.push_up(X_local[s0:s1], mode, self.PS[r], s0, self.center, None, self.stat[own])
                if t1 > t0:
                    ops.pairwise_tiles(self.PS, N, mode, plan.slice(t0, t1), self.flags, self.epoch, D_comp, sh, n * sh, stat,
                                       flist, self.list_cap, plan.seg_slice(t0, t1) if forward else None, seg_done)
                    if forward:
                        self._forward_segments(plan, D_comp, D_target, b, b + 1)
            mark("push_end")
            rest, rest_seg, seg_from = plan.slice(self.offdiag_from, None), plan.seg_slice(self.offdiag_from, None), 4
        self._x_local = X_local
        # the statistics of all samples (the epilogue criterion needs both members of a pair); every rank contributes
        # after its push-up, so its arrival also says: all shards are ready to be pulled
        dist.all_gather_into_tensor(self.stat, self.stat[own])
        ready = ops.record_event()
        ops.stream_wait(self.copy_stream, ready)
        for q in self.pulls:                      # NVLink pulls, one after the other, a flag behind each
            ops.copy_async(self.PS[q], self.peer_PS[q].tensor[q], self.copy_stream)
            ops.stream_write_u32(self.flags, q, self.epoch, self.copy_stream)
        mark("pair_begin")
        if rest:
            ops.pairwise_tiles(self.PS, N, mode, rest, self.flags, self.epoch, D_comp, sh, n * sh, stat, flist,
                               self.list_cap, rest_seg if forward else None, seg_done)
        if forward:
            self._forward_segments(plan, D_comp, D_target, seg_from)
            self._end_segments()
        mark("pair_end")
        # the next step's pulls and push-up must not start before everybody is done with this step's: the all-reduce of
        # the flagged-pair count is that barrier (and tells every rank whether the float64 pass is needed)
        self.flagged.copy_(self.flag_list[:1])
        dist.all_reduce(self.flagged)
        self._pending = True
        mark("refine_end")

    def _compute_peer_band_pulls(self, X_local, D_target, mark, host_src):
        """Host-buffer form of the peer step with BAND-WISE pulls (weighted L1, direct --L2) — experimental, FUF_BAND_PULLS=1.

        Every rank uploads its shard in N_PULL_BANDS bands of whole tiles.  As soon as band s has landed it is pushed up
        and announced: a 4-byte copy of the step's epoch into the ``ready`` word (shard, band) of every rank that pulls
        this shard.  The copy stream of such a rank waits for that word (cuStreamWaitValue32), pulls the band — its columns
        of the operand buffer and its slice of the statistics — over NVLink and raises the arrival flag the tile kernel
        polls.  Step s then launches the tiles that became computable with it: the column band s of the diagonal block, and
        of every shard-pair block the boxes {own band s} x {remote bands <= s} and {own bands < s} x {remote band s}
        (``banded_pull_plan``).  Off-diagonal tiles therefore run while later bands are still crossing PCIe, instead of
        after every rank's complete upload; finished boxes leave for the host as completion segments.

        Submission order (it keeps the scheme free of deadlocks even when streams share a hardware queue): per step first
        the announcement, then the waits and pulls of the copy stream, then the tile launch — no wait is ever queued ahead
        of work needed to satisfy it, here or (by induction over the steps) on a peer."""
        ops, N, mode, r, sh, n, dist = self.ops, self.N, self.mode, self.rank, self.shard, self.ops.n, self.dist
        bp, nb = self.bp, N_PULL_BANDS
        plan = bp["plan"]
        self.epoch += 1
        own = slice(r * sh, (r + 1) * sh)
        stat = self.stat if self.do_refine else None
        flist = self.flag_list if self.do_refine else None
        ops.zero_(self.flag_list[:1])
        D_comp = self._local_matrix(D_target)
        forward = D_comp is not D_target
        if forward:
            self._begin_segments()
        ops.fill_(self.epoch_word, self.epoch)
        # the tile launches of this form wait, running, for band pulls — strided peer copies, which the driver may execute
        # as SM kernels: leave a few SMs without a persistent CTA, or such a copy could never start (seen on two B200s:
        # the kernels' bounded waits trapped; the whole-shard pulls of the other forms are contiguous copy-engine transfers)
        ops.set_sm_limit(-8)
        mark("start")
        events = []
        for s0, s1 in bp["own"]:                   # all uploads are queued up front on the copy stream
            if s1 > s0:
                ops.to_device(host_src[s0:s1], X_local[s0:s1], self.copy_stream)
            events.append(ops.record_event(self.copy_stream))
        ops.stream_wait(None, events[0])
        if r == 0:
            ops.push_up_center(X_local[:bp["own"][0][1]], mode, self.center)
        dist.broadcast(self.center, src=0)
        mark("push_begin")
        for s in range(nb):
            s0, s1 = bp["own"][s]
            c0, c1 = bp["cols"][s]
            ops.stream_wait(None, events[s])
            if s1 > s0:
                ops.push_up(X_local[s0:s1], mode, self.PS[r], s0, self.center, None, self.stat[own])
            for q in self.consumers:               # announce the band (also an empty one: nobody may wait forever)
                ops.copy_async(self.peer_ready[q].tensor[r * nb + s:r * nb + s + 1], self.epoch_word, None)
            for q in self.pulls:
                ops.stream_wait_u32(self.ready, q * nb + s, self.epoch, self.copy_stream)
                ops.copy_cols(self.PS[q], self.peer_PS[q].tensor[q], c0, c1, self.copy_stream)
                ops.copy_async(self.stat[q * sh + c0:q * sh + c1], self.peer_stat[q].tensor[q * sh + c0:q * sh + c1],
                               self.copy_stream)
                ops.stream_write_u32(self.band_flags, q * nb + s, self.epoch, self.copy_stream)
            a, b = bp["step_range"][s]
            if b > a:
                ops.pairwise_tiles(self.PS, N, mode, plan.slice(a, b), self.band_flags, self.epoch, D_comp, sh, n * sh, stat,
                                   flist, self.list_cap, plan.seg_slice(a, b) if forward else None,
                                   self.seg_done if forward else None)
                if forward:
                    self._forward_segments(plan, D_comp, D_target, *bp["seg_range"][s])
        ops.set_sm_limit(0)
        mark("push_end")
        self._x_local = X_local
        mark("pair_begin")
        if forward:
            self._end_segments()
        mark("pair_end")
        # the next step's pulls and push-ups must not start before everybody is done with this step's (see _compute_peer)
        self.flagged.copy_(self.flag_list[:1])
        dist.all_reduce(self.flagged)
        self._pending = True
        mark("refine_end")

    def _refine_peer(self, D_target) -> None:
        """The host's half of the refinement decision (one synchronisation per step; the float64 pass itself is rare)."""
        self.last_flagged = 0
        if not self._pending:
            return
        self._pending = None
        self.last_flagged = int(self.flagged.item())
        if self.last_flagged == 0 or not self.do_refine:
            return
        ops, r, sh, n, dist = self.ops, self.rank, self.shard, self.ops.n, self.dist
        if self.P64S is None:
            self.P64S = ops.empty((self.world, n, sh), "float64")
        if self.hi > self.lo:
            ops.push_up_p64(self._x_local, self.mode, self.P64S[r])
        dist.all_gather_into_tensor(self.P64S.view(-1), self.P64S[r].reshape(-1))
        counts = ops.empty((self.world,), "int64")
        dist.all_gather_into_tensor(counts, self.flag_list[:1].clone())
        if int(counts.max().item()) > self.list_cap:
            # more flagged pairs than a list holds (a data set of near-duplicates): the float64 tile kernel redoes the
            # whole matrix on rank 0, which holds every float64 profile now
            if r == 0:
                full = self.D if D_target is self.D_target else ops.empty((self.N, self.N), "float64")
                ops.pairwise_f64(self.P64S, self.N, self.mode, None, sh, n * sh, full)
                if full is not self.D:
                    ops.copy_to_target(full, D_target)
        else:
            ops.refine_list(self.P64S, self.N, self.mode, self.flag_list, self.list_cap, D_target, sh, n * sh)
        ops.synchronize()
        dist.barrier()                            # every rank's corrections have landed

    def step(self, X_local, mark=None):
        """X_local: this rank's un-pushed samples [hi - lo, >= n] on the device.  Returns D on rank 0, else None.
        ``mark(name)`` is called at the phase boundaries (bench.py records CUDA events there)."""
        mark = mark or (lambda name: None)
        if self.use_peer:
            self._compute_peer(X_local, self.D_target, mark)
            self._refine_peer(self.D_target)
            mark("end")
            return self.D
        self._compute(X_local, mark)
        if self.world == 1:
            self._refine_if_flagged()
        else:
            r = self.rank
            if self.packed is None:
                self.packed = self.ops.empty((self.max_packed,), "float64")
                if r == 0:
                    self.D = self.ops.empty((self.N, self.N), "float64")
                    self.gathered = self.ops.empty((self.world, self.max_packed), "float64")
            for attempt in range(2):
                if self.rows:
                    self.ops.pack_tile_rows(self.blocks, self.rows, self.offs, self.N, self.packed)
                self.dist.gather(self.packed, [self.gathered[q] for q in range(self.world)] if r == 0 else None, dst=0)
                if attempt == 0:
                    mark("gather_end")
                if r == 0:
                    self.ops.assemble_packed(self.gathered, self.slot_rows, self.slot_offs, self.N, self.D)
                if attempt == 1 or not self._refine_if_flagged():
                    break
        mark("end")
        return self.D

    def step_host(self, X_local_host, D_host: "SharedHostMatrix", mark=None):
        """Host buffers on both sides: ``X_local_host`` is this rank's un-pushed samples in (pinned) host memory,
        ``D_host`` the node-wide shared matrix.  Every rank copies its samples in, computes its tile rows and stores
        them and their mirror image into ``D_host``; when the call returns (after a barrier) the full symmetric
        matrix is in ``D_host.array`` on EVERY rank."""
        mark = mark or (lambda name: None)
        if D_host.array.shape != (self.N, self.N):
            raise ValueError(f"shared matrix is {D_host.array.shape}, job needs {(self.N, self.N)}")
        if self._x_stage is None:
            self._x_stage = self.ops.profile_matrix(max(self.hi - self.lo, 1))
        if self.use_peer:
            # every finished completion segment of the rank's tiles (and their mirror) is copied by a DMA engine into the
            # page-locked host matrix the ranks share: the device-to-host traffic of all ranks runs over their own
            # PCIe links, under the computation
            target = self.ops.host_target(D_host)
            self._compute_peer(self._x_stage[:self.hi - self.lo], target, mark, host_src=X_local_host)
            self._refine_peer(target)
            self.ops.synchronize()
            self.dist.barrier()
            mark("end")
            return D_host.array
        X = self.ops.to_device(X_local_host, self._x_stage)
        self._compute(X, mark)
        if self.world == 1:
            self._refine_if_flagged()
            self.ops.copy_to_host(self.D, D_host.array)
        else:
            for attempt in range(2):
                if self.rows:
                    self.ops.store_tile_rows_host(self.blocks, self.rows, self.N, D_host.array)
                if attempt == 1 or not self._refine_if_flagged():
                    break
        self.ops.synchronize()
        if self.world > 1:
            self.dist.barrier()
        mark("end")
        return D_host.array


class _SegmentPlan:
    """Work list of a rank (tiles, or Gram items) with the completion segment of every entry, the rectangles each segment
    finishes and the counter value that says so.  Slices are cached so that the ops layer converts each list once."""

    def __init__(self, work, seg, regions, units: int = 1):
        self.work, self.seg, self.regions = work, seg, regions
        self.n_seg = len(regions)
        self.expected = [0] * self.n_seg
        for sg in seg:
            self.expected[sg] += units
        self._slices = {}

    def slice(self, a, b):
        key = ("w", a, b)
        if key not in self._slices:
            self._slices[key] = self.work[a:b]
        return self._slices[key]

    def seg_slice(self, a, b):
        key = ("s", a, b)
        if key not in self._slices:
            self._slices[key] = self.seg[a:b]
        return self._slices[key]


class SharedHostMatrix:
    """N x N float64 host matrix shared by the ranks of ONE node, page-locked for every rank's GPU.

    Rank 0 creates an anonymous memory file (``memfd_create``: no name, no ``/dev/shm`` size limit, freed when the last
    process closes it); the other ranks open it through ``/proc/<pid>/fd/<fd>`` and map the same pages.  Each rank
    registers its mapping with CUDA (``cudaHostRegister``), so ``fuf_store_tile_rows_host`` DMA-writes into it over
    that rank's own PCIe link.  This is the array the driver then hands to ``np.save``
    (``src/compute_all_pairwise_fununifrac.py:88-89``)."""

    def __init__(self, N: int, world: int = 1, rank: int = 0, dist=None, register: Optional[bool] = None):
        import mmap
        import os

        import numpy as np

        self.N, self.nbytes = N, max(8, N * N * 8)
        self._registered = False
        self._fd = self._map = None
        if rank == 0:
            self._fd = os.memfd_create("fununifrac_D")
            os.ftruncate(self._fd, self.nbytes)
        if world > 1:
            where = [(os.getpid(), self._fd) if rank == 0 else None]
            dist.broadcast_object_list(where, src=0)
            opened = True
            if rank != 0:
                try:
                    if os.environ.get("FUF_SHARED_NO_PROCFD"):      # test hook: exercise the named-file route
                        raise OSError("disabled")
                    self._fd = os.open(f"/proc/{where[0][0]}/fd/{where[0][1]}", os.O_RDWR)
                except OSError:
                    opened = False
            votes = [None] * world
            dist.all_gather_object(votes, opened)      # rank 0 keeps its descriptor open until everybody has answered
            if not all(votes):
                # /proc/<pid>/fd is not reachable from every rank (separate PID namespaces): a named file in shared memory
                import tempfile
                for fd in ([self._fd] if self._fd is not None else []):
                    os.close(fd)
                path = [None]
                if rank == 0:
                    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else None
                    self._fd, path[0] = tempfile.mkstemp(prefix="fununifrac_D_", dir=shm_dir)
                    os.ftruncate(self._fd, self.nbytes)
                dist.broadcast_object_list(path, src=0)
                if rank != 0:
                    self._fd = os.open(path[0], os.O_RDWR)
                dist.barrier()
                if rank == 0:
                    os.unlink(path[0])                 # the pages live until the last mapping goes
        self._map = mmap.mmap(self._fd, self.nbytes)
        self.array = np.frombuffer(self._map, dtype=np.float64, count=N * N).reshape(N, N)
        if register is None:
            import torch
            register = torch.cuda.is_available()
        if register and N:
            import torch
            import ctypes

            from fununifrac_b200 import engine
            torch.cuda.init()
            pieces = ctypes.c_int32()
            engine._check(engine.load_library().fuf_host_register(self.array.ctypes.data, self.nbytes, ctypes.byref(pieces)),
                          "fuf_host_register")
            self._pieces = pieces.value
            self._registered = True

    def close(self):
        import os
        if self._registered:
            from fununifrac_b200 import engine
            engine.load_library().fuf_host_unregister(self.array.ctypes.data, self.nbytes, self._pieces)
            self._registered = False
        self.array = None
        if self._map is not None:
            try:
                self._map.close()
            except BufferError:      # a view of the array is still alive somewhere: the pages go when it does
                pass
            self._map = None
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None


class CudaOps:
    """``ops`` over ``engine.TreeContext`` (the only implementation the product uses)."""

    def __init__(self, ctx):
        import torch

        self.ctx, self.n, self.torch = ctx, ctx.n, torch
        self.device = f"cuda:{ctx.device}"

    def empty(self, shape, dtype):
        return self.torch.zeros(shape, dtype=getattr(self.torch, dtype), device=self.device)

    def push_up_center(self, X, mode, out):
        out.copy_(self.ctx.push_up_center(X, mode))

    def push_up(self, X, mode, PT, col0, center, P64T, err_stat):
        return self.ctx.push_up(X, mode, PT, col0, center, P64T, err_stat)

    def profile_matrix(self, rows):
        from fununifrac_b200 import engine
        return engine.profile_matrix(rows, self.n, self.torch.float64, self.device)

    def peer_alloc(self, shape, dtype):
        return self.ctx.peer_alloc(shape, dtype)

    def peer_open(self, handle, shape, dtype):
        return self.ctx.peer_open(handle, shape, dtype)

    def new_stream(self):
        return self.torch.cuda.Stream(device=self.device)

    def record_event(self, stream=None):
        ev = self.torch.cuda.Event()
        ev.record(stream if stream is not None else self.torch.cuda.current_stream())
        return ev

    def stream_wait(self, stream, event):
        (stream if stream is not None else self.torch.cuda.current_stream()).wait_event(event)

    def zero_(self, t):
        t.zero_()

    def copy_async(self, dst, src, stream):
        self.ctx.copy_async(dst, src, stream)

    def stream_write_u32(self, flags, index, value, stream):
        self.ctx.stream_write_u32(flags, index, value, stream)

    gram_seg_units = 16          # engine.GRAM_SEG_UNITS: counts per finished item of fuf_gram_tiles

    def _as_np(self, lst, width):
        """The plans are fixed for the job: every list is converted to a numpy array once."""
        cache = self.__dict__.setdefault("_lists_np", {})
        arr = cache.get(id(lst))
        if arr is None or arr[0] is not lst:
            import numpy as np
            a = np.ascontiguousarray(lst, dtype=np.int32)
            arr = cache[id(lst)] = (lst, a.reshape(-1, width) if width > 1 else a.reshape(-1))
        return arr[1]

    def pairwise_tiles(self, PS, N, mode, tiles, flags, epoch, D, shard, shard_stride, stat, flag_list, list_cap,
                       seg=None, seg_done=None):
        self.ctx.pairwise_tiles(PS, N, mode, self._as_np(tiles, 3), flags, epoch, D, shard, shard_stride, stat, flag_list,
                                list_cap, segments=None if seg is None else self._as_np(seg, 1), seg_done=seg_done)

    def stream_wait_u32(self, words, index, value, stream, exact=False):
        self.ctx.stream_wait_u32(words, index, value, stream)

    def fill_(self, t, value):
        t.fill_(value)

    def set_sm_limit(self, sms):
        """sms > 0: size the persistent grids for that many SMs; sms < 0: leave -sms SMs free; 0: the context's default."""
        self.ctx.set_sm_limit(sms)

    def copy_cols(self, dst, src, c0, c1, stream):
        """dst[:, c0:c1] = src[:, c0:c1] for two [rows, cols] buffers of the same layout (a band of a shard's operands,
        pulled from the peer that owns it): one pitched copy-engine transfer."""
        self.ctx.copy_cols_async(dst, src, c0, c1, stream)

    def copy_block(self, target, src, rows, cols, stream):
        """target[rows, cols] = src[rows, cols]; ``target`` is an N x N tensor (peer mapping) or the raw device-side pointer
        of the shared host matrix (row pitch N)."""
        ld = src.shape[1] if isinstance(target, int) else target.stride(0)
        self.ctx.copy_block_async(target, ld, src, rows, cols, stream)

    def gram_colstats(self, PT_shard, n_valid, stats):
        self.ctx.gram_colstats(PT_shard, n_valid, stats)

    def gram_plan(self, stats):
        return self.ctx.gram_plan(stats)

    def gram_digits(self, PT_shard, n_valid, planes, shard_index, err_in, q, h2, err_out, SUB):
        self.ctx.gram_digits(PT_shard, n_valid, planes, shard_index, err_in, q, h2, err_out, SUB)

    def gram_tiles(self, planes, N, items, flags, epoch, q, h2, err, SUB, D, flag_list, list_cap, seg=None, seg_done=None):
        self.ctx.gram_tiles(planes, N, self._as_np(items, 3), flags, epoch, q, h2, err, SUB, D, flag_list, list_cap,
                            segments=None if seg is None else self._as_np(seg, 1), seg_done=seg_done)

    def refine_list(self, P64S, N, mode, flag_list, list_cap, D, shard, shard_stride):
        self.ctx.refine_list(P64S, N, mode, flag_list, list_cap, D, shard, shard_stride)

    def host_target(self, D_host):
        """Raw device-side pointer of the shared page-locked host matrix (the kernels store through it)."""
        if getattr(self, "_host_target_of", None) is not D_host:
            self._host_target, self._host_target_of = self.ctx.host_device_pointer(D_host.array), D_host
        return self._host_target

    def copy_to_target(self, D, target):
        if isinstance(target, int):
            import ctypes
            nbytes = D.numel() * 8
            self.ctx._lib.fuf_copy_async(self.ctx.handle, ctypes.c_void_p(target), D.data_ptr(), nbytes,
                                         self.torch.cuda.current_stream().cuda_stream)
        else:
            target.copy_(D)

    def push_up_p64(self, X, mode, P64T):
        return self.ctx.push_up_p64(X, mode, P64T)

    def pairwise(self, PT, N, mode, tile_rows, shard, shard_stride, D):
        return self.ctx.pairwise(PT, N, mode, D=D, tile_rows=tile_rows, shard=shard, shard_stride=shard_stride)

    def pairwise_gram(self, PT, N, err_stat, tile_rows, shard, shard_stride, D):
        _, q, err = self.ctx.pairwise_gram(PT, N, D=D, err_stat=err_stat, tile_rows=tile_rows, shard=shard,
                                           shard_stride=shard_stride)
        return q, err

    def refine(self, P64T, N, mode, err_stat, D, tile_rows, shard, shard_stride, lin_stat=None):
        return self.ctx.refine(P64T, N, mode, err_stat, D, tile_rows=tile_rows, shard=shard, shard_stride=shard_stride,
                               lin_stat=lin_stat, lin_kappa=2.9e-3)

    def pairwise_f64(self, P64T, N, mode, tile_rows, shard, shard_stride, D):
        return self.ctx.pairwise_f64(P64T, N, mode, D=D, tile_rows=tile_rows, shard=shard, shard_stride=shard_stride)

    def count_uncertified_into(self, N, mode, err_stat, D, tile_rows, lin_stat, out):
        self.ctx.refine(None, N, mode, err_stat, D, tile_rows=tile_rows, lin_stat=lin_stat, lin_kappa=2.9e-3)
        self.ctx.refine_count_into(out)

    def assemble(self, blocks, rows, N, D):
        return self.ctx.assemble(blocks, rows, N, D=D)

    def pack_tile_rows(self, blocks, rows, offsets, N, packed):
        return self.ctx.pack_tile_rows(blocks, rows, offsets, N, packed)

    def assemble_packed(self, packed, rows, offsets, N, D):
        return self.ctx.assemble_packed(packed.view(-1), rows, offsets, N, D=D)

    def to_device(self, X_host, stage, stream=None):
        """Asynchronous host-to-device copy of (page-locked) rows into ``stage`` on ``stream`` (default: current)."""
        src = X_host if self.torch.is_tensor(X_host) else self.torch.from_numpy(X_host)
        out = stage[:src.shape[0]]
        if stream is None:
            out.copy_(src, non_blocking=True)
        else:
            with self.torch.cuda.stream(stream):
                out.copy_(src, non_blocking=True)
        return out

    def store_tile_rows_host(self, blocks, rows, N, D_host):
        self.ctx.store_tile_rows_host(blocks, rows, N, D_host)

    def copy_to_host(self, D, D_host):
        self.torch.from_numpy(D_host).copy_(D, non_blocking=True)

    def synchronize(self):
        self.torch.cuda.synchronize()
