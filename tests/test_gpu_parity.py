"""GPU parity tests (run with ``-m gpu`` on the B200 box): the CUDA path, called through the C ABI, against
(1) vectors produced by the unmodified reference (tests/golden), (2) the CPU oracle on seeded inputs, and
(3) size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star): tree/postorder/basis bit-exact (CPU tests); float64 validation
kernels: push-up bit-identical, distances within 1e-12 relative; production path (centred fp32 operands +
float64 refinement of uncertified pairs): within 1e-6 relative of the reference's float64 output for EVERY pair,
near-duplicates included, and exactly 0 for identical samples.
"""
import glob
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch  # noqa: E402

from fununifrac_b200 import engine, sparse_npz, synth  # noqa: E402
from fununifrac_b200.algorithms.emd_unifrac import EarthMoverDistanceUniFracSolver  # noqa: E402
from fununifrac_b200.factory import make_emd_input, make_tree  # noqa: E402
from oracle import oracle_c  # noqa: E402

REL = 1e-6


def rel_err(D, ref, scale=None):
    iu = np.triu_indices(len(ref), 1)
    den = np.maximum(ref[iu], 1e-300) if scale is None else np.maximum(ref[iu], scale[iu])
    return (np.abs(D[iu] - ref[iu]) / den).max() if len(iu[0]) else 0.0


def run_all(ctx, X, l2, check_scalar=True):
    """Every device route for the same input: validation, production (packed / scalar), host-buffer call."""
    mode = engine.MODE_L2 if l2 else engine.MODE_L1
    N = X.shape[0]
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    P64T = ctx.push_up_f64(Xd, mode)
    D64 = ctx.pairwise_f64(P64T, N, mode).cpu().numpy()
    PT = ctx.push_up(Xd, mode)                      # uncentred fp32 profiles: checked against the reference P
    job = ctx.push_up_job(Xd, mode)                 # production: centred fp32 + float64 copy + error statistics
    # the production float64 copy: the reference push-up up to the rounding of its own additions (nodes whose
    # subtree crosses a 32-node chunk are summed chunk-wise instead of child by child)
    assert torch.allclose(job.P64T[:, :N], P64T[:, :N], rtol=1e-13, atol=0)
    D32 = ctx.distances(job, mode).cpu().numpy()
    if check_scalar:
        Draw = ctx.pairwise(job.PT, N, mode)
        Draw_s = ctx.pairwise(job.PT, N, mode, flags=engine.PW_SCALAR_MATH)
        assert torch.equal(Draw, Draw_s)            # FADD2/FFMA2 are IEEE per lane: identical to scalar FADD/FFMA
    Dh = ctx.all_pairs_host(X, mode)
    assert np.array_equal(Dh, D32)
    Dhv = ctx.all_pairs_host(X, mode, validate=True)
    assert np.array_equal(Dhv, D64)
    return P64T[:, :N].t().cpu().numpy(), PT[:, :N].t().cpu().numpy(), D64, D32


def check_matrix(D):
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0) and np.all(D >= 0)


# --------------------------------------------------------------------------------------------------
# (1) reference-generated goldens
# --------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def toy(data_dir, golden_dir):
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"))
    inp = make_emd_input.tree_to_EMDU_input(tree, "ko00001")
    return inp, engine.TreeContext(inp.parent, inp.edge_length), np.load(os.path.join(golden_dir, "toy_config1.npz"))


@pytest.mark.parametrize("tag,l2", [("l1_median", False), ("l2_median", True), ("l1_median_unweighted", False),
                                    ("l2_median_unweighted", True), ("l1_default", False)])
def test_config1_toy(tag, l2, toy):
    inp, ctx, z = toy
    P64, P32, D64, D32 = run_all(ctx, z[f"{tag}.X"], l2)
    assert np.array_equal(P64, z[f"{tag}.P"])                      # bit-identical push-up
    assert np.allclose(P32, z[f"{tag}.P"], rtol=2e-7, atol=0)
    ref = z[f"{tag}.D"]
    for D in (D64, D32):
        check_matrix(D)
    assert rel_err(D64, ref) < 1e-12 and rel_err(D32, ref) < REL
    assert np.all(D32[ref == 0] == 0)                              # identical samples → exactly 0
    if tag == "l1_median":   # hand values of tests/scripts/test_make_all_pw_fununifrac.py:32-33
        assert np.allclose(D32, [[0.0, 13 / 14, 10 / 21], [13 / 14, 0.0, 4 / 3], [10 / 21, 4 / 3, 0.0]], atol=1e-3)
    if tag == "l2_median":   # :173-175
        k = np.array([[0.0, np.sqrt(37) / 14, np.sqrt(22) / 21], [np.sqrt(37) / 14, 0.0, np.sqrt(13) / 6],
                      [np.sqrt(22) / 21, np.sqrt(13) / 6, 0.0]]) ** 2
        assert np.allclose(D32, k, atol=1e-3)


@pytest.mark.parametrize("tag,l2,unw", [("l1", False, False), ("l2", True, False), ("l1_unweighted", False, True)])
def test_synth_mid_golden(tag, l2, unw, golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_mid.npz"))
    ctx = engine.TreeContext(z["parent"], z["length"])
    assert ctx.info().n_zero_len > 0 and ctx.info().n_spanning > 0
    X = z["X"].copy()
    if unw:
        X[X > 0] = 1
    P64, P32, D64, D32 = run_all(ctx, X, l2)
    assert np.array_equal(P64, z[f"{tag}.P"])
    assert np.allclose(P32, z[f"{tag}.P"], rtol=2e-7, atol=0)
    ref = z[f"{tag}.D"]
    check_matrix(D32)
    assert rel_err(D64, ref) < 1e-12
    # every pair within 1e-6 relative: ordinary ones from the fp32 tiles, the near-duplicates (0,20) through the
    # float64 refinement pass; the identical pair (1,21) is exactly 0
    assert rel_err(D32, ref) < REL
    assert D32[1, 21] == 0.0 and D64[1, 21] == 0.0


def test_real_tree_golden(golden_dir, real_trees):
    z = np.load(os.path.join(golden_dir, "real_tree_rows.npz"))
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(
        real_trees["kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt"]), "ko00001")
    ctx = engine.TreeContext(inp.parent, inp.edge_length)
    n, N = len(inp.basis), len(z["indptr"]) - 1
    X = np.zeros((N, n))
    for i in range(N):
        s, e = z["indptr"][i], z["indptr"][i + 1]
        X[i, z["indices"][s:e]] = z["values"][s:e]
    internal = np.setdiff1d(np.arange(n), synth.leaves_of(inp.parent))
    for tag, l2 in (("l1", False), ("l2", True)):
        P64, P32, D64, D32 = run_all(ctx, X, l2)
        assert np.array_equal(P64[0][internal], z[f"{tag}.P0_internal"])
        assert rel_err(D64, z[f"{tag}.D"]) < 1e-12 and rel_err(D32, z[f"{tag}.D"]) < REL


# --------------------------------------------------------------------------------------------------
# (2) the oracle on seeded inputs, full-size synthetic KO tree
# --------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_tree(tmp_path_factory):
    path = synth.synth_ko_tree(str(tmp_path_factory.mktemp("synth") / "ko_tree.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
    assert len(inp.basis) == 30002
    return inp, engine.TreeContext(inp.parent, inp.edge_length), synth.leaves_of(inp.parent)


@pytest.mark.parametrize("l2", [False, True])
def test_full_tree_vs_oracle(l2, full_tree):
    inp, ctx, leaves = full_tree
    N = 300                                  # 3 tile rows, edge tile 44 wide, k-split path (few tiles)
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=1)
    Pref = oracle_c.push_up(inp.parent, inp.edge_length, X, l2)
    Dref = oracle_c.pairwise(Pref, l2)
    P64, P32, D64, D32 = run_all(ctx, X, l2, check_scalar=False)
    assert np.array_equal(P64, Pref)
    assert np.allclose(P32, Pref, rtol=2e-7, atol=0)
    check_matrix(D32)
    assert rel_err(D64, Dref) < 1e-12
    assert rel_err(D32, Dref) < REL
    # float32 profiles as INPUT (a caller's choice, not ours): the float32 values are the data, so the oracle
    # gets the same rounded profiles
    mode = engine.MODE_L2 if l2 else engine.MODE_L1
    X32 = X.astype(np.float32)
    D = ctx.distances(ctx.push_up_job(torch.from_numpy(X32).cuda(), mode), mode).cpu().numpy()
    Dref32 = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X32.astype(np.float64), l2), l2)
    assert rel_err(D, Dref32) < REL


@pytest.mark.parametrize("l2", [False, True])
def test_tma_pushup_kernel_vs_chunk_kernel_and_oracle(l2, full_tree, monkeypatch):
    """The production push-up (pushup_tma_kernel: TMA boxes of X, prefix differences inside a chunk) against the chunk
    kernel (child-by-child additions) and the oracle: float64 and float32 input, N that is not a multiple of the 256-sample
    tile, a column offset, a padded row pitch, and the per-sample rounding statistic."""
    inp, ctx, leaves = full_tree
    mode = engine.MODE_L2 if l2 else engine.MODE_L1
    N, n = 700, ctx.n
    X = synth.synth_profiles_dense(leaves, n, N, seed=21)
    X[:, -1] = 1e-3                                   # mass on internal nodes too (root, an upper-level node)
    X[:, -2] = 2e-3
    Pref = oracle_c.push_up(inp.parent, inp.edge_length, X, l2)
    Xd = torch.from_numpy(X).cuda()
    center = ctx.push_up_center(Xd, mode)
    cen = center.cpu().numpy()
    want = Pref - cen[None, :]

    def run(Xin, col0=0):
        PT = ctx.new_profile_buffer(N + col0)
        stat = torch.zeros(PT.shape[1], dtype=torch.float64, device="cuda")
        ctx.push_up(Xin, mode, PT, col0, center, None, stat)
        return PT[:, col0:col0 + N].t().cpu().numpy(), stat[col0:col0 + N].cpu().numpy()

    launches = ctx.launch_count()
    P_tma, s_tma = run(Xd)
    n_tma = ctx.launch_count() - launches
    monkeypatch.setenv("FUF_PUSHUP_NO_TMA", "1")
    P_chunk, s_chunk = run(Xd)
    monkeypatch.delenv("FUF_PUSHUP_NO_TMA")
    # both round the same float64 value up to ~1e-16 of the chunk's mass: identical except for rare 1-ulp ties
    assert np.allclose(P_tma, want, rtol=2e-7, atol=1e-7 * np.abs(want).max() * 1e-3)
    ulp = np.spacing(np.abs(P_chunk).astype(np.float32)).astype(np.float64)
    # (+ the float64 noise of the sums themselves where the centred value cancels to ~0: root and upper levels)
    assert np.all(np.abs(P_tma.astype(np.float64) - P_chunk) <= ulp + 1e-14 * np.abs(Pref).max())
    assert (P_tma != P_chunk).mean() < 1e-3
    resid = want - P_tma.astype(np.float64)
    stat_true = (resid ** 2).sum(1) if l2 else np.abs(resid).sum(1)
    assert np.allclose(s_tma, stat_true, rtol=1e-3) and np.allclose(s_tma, s_chunk, rtol=1e-3)
    assert np.array_equal(P_tma[:, leaves] == 0, X[:, leaves] == 0)          # exact zeros stay exact zeros (leaves: no centre)
    # column offset and a padded, 16-byte aligned row pitch (what fuf_all_pairs_host uses for odd n)
    P_off, s_off = run(Xd, col0=128)
    assert np.array_equal(P_off, P_tma) and np.array_equal(s_off, s_tma)
    Xpad = torch.zeros((N, n + 6), dtype=torch.float64, device="cuda")
    Xpad[:, :n] = Xd
    P_pad, s_pad = run(Xpad[:, :n])
    assert np.array_equal(P_pad, P_tma) and np.array_equal(s_pad, s_tma)
    # an unaligned pitch silently takes the chunk kernel: same values as the chunk kernel
    Xodd = torch.zeros((N, n + 1), dtype=torch.float64, device="cuda")
    Xodd[:, :n] = Xd
    P_odd, _ = run(Xodd[:, :n])
    assert np.array_equal(P_odd, P_chunk)
    # float32 input
    X32 = Xd.float()
    center32 = ctx.push_up_center(X32, mode)
    PT32 = ctx.new_profile_buffer(N)
    ctx.push_up(X32, mode, PT32, 0, center32)
    want32 = oracle_c.push_up(inp.parent, inp.edge_length, X32.double().cpu().numpy(), l2) - center32.cpu().numpy()[None, :]
    assert np.allclose(PT32[:, :N].t().cpu().numpy(), want32, rtol=2e-7, atol=1e-10 * np.abs(want32).max())
    assert n_tma >= 1
    # the same kernel on the real KO tree (odd n = 25,961: padded pitch) is covered by test_real_tree_golden through
    # fuf_all_pairs_host


def test_unweighted_full_tree(full_tree):
    inp, ctx, leaves = full_tree
    X = synth.synth_profiles_dense(leaves, ctx.n, 40, seed=2)
    X[X > 0] = 1                              # --unweighted: after normalisation, not renormalised
    Pref = oracle_c.push_up(inp.parent, inp.edge_length, X, False)
    assert Pref[:, -1].min() > 1000           # the unscaled root entry carries the node count
    Dref = oracle_c.pairwise(Pref, False)
    D = ctx.all_pairs_host(X, engine.MODE_L1)
    assert rel_err(D, Dref) < REL


# --------------------------------------------------------------------------------------------------
# sharding primitives: compact tile rows + assemble, sharded profile layout
# --------------------------------------------------------------------------------------------------
def test_tile_rows_assemble_and_shards(full_tree):
    inp, ctx, leaves = full_tree
    N = 700                                   # 6 tile rows
    X = torch.from_numpy(synth.synth_profiles_dense(leaves, ctx.n, N, seed=3, dtype=np.float32)).cuda()
    X[5] = X[4] * (1 + 1e-4 * torch.rand(ctx.n, device="cuda"))       # a near-duplicate pair inside one tile
    X[650] = X[3]                                                     # an identical pair across tiles
    job = ctx.push_up_job(X, engine.MODE_L1)
    D = ctx.distances(job, engine.MODE_L1)
    assert ctx.refine_count() >= 2 and D[3, 650] == 0 and D[650, 3] == 0
    T = (N + 127) // 128
    # 3 "ranks", boustrophedon tile rows, each computing (and refining) compact upper row blocks
    owners = [[t for t in range(T) if (t % 6 if t % 6 < 3 else 5 - t % 6) == r] for r in range(3)]
    blocks, rows = [], []
    for own in owners:
        blk = ctx.pairwise(job.PT, N, engine.MODE_L1, tile_rows=own)
        assert blk.shape == (len(own) * 128, N)
        ctx.refine(job.P64T, N, engine.MODE_L1, job.err_stat, blk, tile_rows=own)
        blocks.append(blk)
        rows += own
    D2 = ctx.assemble(torch.cat(blocks, 0), rows, N)
    # the number of k-splits depends on the tile count: same fp32 partial sums, different order of the final
    # float64 additions
    bad = (D - D2).abs() > 1e-13 * D.abs()
    assert not bool(bad.any()), (bad.nonzero()[:8].tolist(), D[bad][:8].tolist(), D2[bad][:8].tolist())
    # padding slots (-1) of a gathered buffer are skipped
    D2b = ctx.assemble(torch.cat(blocks + [torch.full((128, N), 7.0, dtype=torch.float64, device="cuda")], 0),
                       rows + [-1], N)
    assert torch.equal(D2, D2b)
    # the same join straight into one host matrix (what every rank of a node does in ShardedAllPairs.step_host):
    # pinned, page-locked-by-registration (SharedHostMatrix) and plain pageable destinations, N not a multiple of 128
    from fununifrac_b200 import sharded
    # packed row blocks (what crosses NVLink in ShardedAllPairs.step): same matrix, bit for bit
    offs_all, packs, base = [], [], 0
    for own, blk in zip(owners, blocks):
        offs, total = sharded.packed_offsets(own, N)
        assert total == sum(128 * (N - 128 * t) for t in own)
        pk = torch.full((total + 5,), float("nan"), dtype=torch.float64, device="cuda")
        ctx.pack_tile_rows(blk, own, offs, N, pk)
        assert bool(pk[total:].isnan().all())            # nothing beyond the slots is touched
        packs.append(pk)
        offs_all += [base + o for o in offs]
        base += total + 5
    D2c = ctx.assemble_packed(torch.cat(packs), rows + [-1], offs_all + [0], N)
    assert torch.equal(D2, D2c)
    shared = sharded.SharedHostMatrix(N)
    pinned = torch.full((N, N), float("nan"), dtype=torch.float64, pin_memory=True).numpy()
    pageable = np.full((N, N + 3), np.nan)[:, :N]
    for dst in (shared.array, pinned, pageable):
        dst[...] = np.nan
        for own, blk in zip(owners, blocks):
            ctx.store_tile_rows_host(blk.clone(), own, N, dst)
        torch.cuda.synchronize()
        assert np.array_equal(dst, D2.cpu().numpy())
    shared.close()
    # sharded layout [G, n, shard]: what an all-gather of per-rank push-up outputs looks like
    shard = 256
    G = (N + shard - 1) // shard
    PS = torch.zeros((G, ctx.n, shard), dtype=torch.float32, device="cuda")
    P64S = torch.zeros((G, ctx.n, shard), dtype=torch.float64, device="cuda")
    stat = torch.zeros(G * shard, dtype=torch.float64, device="cuda")
    for g in range(G):
        w = min(shard, N - g * shard)
        ctx.push_up(X[g * shard:g * shard + w], engine.MODE_L1, PS[g], 0, job.center, None,
                    stat[g * shard:(g + 1) * shard])
        ctx.push_up_p64(X[g * shard:g * shard + w], engine.MODE_L1, P64S[g])
    assert torch.equal(stat[:N], job.err_stat[:N])
    D3 = ctx.pairwise(PS, N, engine.MODE_L1, shard=shard, shard_stride=ctx.n * shard)
    ctx.refine(P64S, N, engine.MODE_L1, stat, D3, shard=shard, shard_stride=ctx.n * shard)
    assert torch.equal(D, D3)
    # the same through the sharded driver at world size 1
    one = sharded.ShardedAllPairs(sharded.CudaOps(ctx), N, engine.MODE_L1)
    assert torch.equal(one.step(X), D)
    shared = sharded.SharedHostMatrix(N)
    assert np.array_equal(one.step_host(X.double().cpu().numpy(), shared), D.cpu().numpy())
    shared.close()


def test_host_path_bands(full_tree, monkeypatch):
    """fuf_all_pairs_host computes D band by band while later samples are still being copied; forcing 3 and 5
    bands at a small N must give the device-resident result up to the order of the final float64 additions."""
    inp, ctx, leaves = full_tree
    N = 700
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=9)
    X[300] = X[2] * (1 + 1e-5 * np.random.default_rng(1).random(ctx.n))
    X[699] = X[130]
    D = ctx.distances(ctx.push_up_job(torch.from_numpy(X).cuda(), engine.MODE_L1), engine.MODE_L1).cpu().numpy()
    for bands in ("1", "3", "5"):
        monkeypatch.setenv("FUF_HOST_BANDS", bands)
        Dh = ctx.all_pairs_host(X, engine.MODE_L1)
        assert np.allclose(Dh, D, rtol=1e-13, atol=0) and np.array_equal(Dh, Dh.T)
        assert Dh[130, 699] == 0 and ctx.refine_count() >= 2
        Dp = engine.pinned_empty((N, N), torch.float64)[1]
        Dp[:] = -1
        ctx.all_pairs_host(X, engine.MODE_L1, Dp)            # pinned destination: written by the tile epilogues
        assert np.array_equal(Dp, Dh)


def test_near_duplicates_are_refined(full_tree):
    """Clusters of near-identical samples: the fp32 operands cannot resolve their distances, the refinement pass
    does; ordinary pairs are left to the fp32 tiles."""
    inp, ctx, leaves = full_tree
    N = 260
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=7)
    rng = np.random.default_rng(7)
    for base, copies, jitter in ((0, 6, 1e-3), (10, 5, 1e-5), (130, 4, 1e-7), (200, 3, 0.0)):
        for c in range(1, copies):
            X[base + c] = X[base] * (1 + jitter * rng.random(ctx.n))
            X[base + c] /= X[base + c].sum()
    for l2 in (False, True):
        mode = engine.MODE_L2 if l2 else engine.MODE_L1
        Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, l2), l2)
        job = ctx.push_up_job(torch.from_numpy(X).cuda(), mode)
        raw = ctx.pairwise(job.PT, N, mode).cpu().numpy()
        D = ctx.distances(job, mode).cpu().numpy()
        n_ref = ctx.refine_count()
        assert rel_err(D, Dref) < REL and np.array_equal(D, D.T)
        assert rel_err(raw, Dref) > REL                      # without the pass the bar is missed ...
        assert 15 <= n_ref <= 400                            # ... and only the near-duplicates pay for it
        assert D[200, 201] == 0 and D[201, 202] == 0
        Dh = ctx.all_pairs_host(X, mode)
        assert np.array_equal(Dh, D)
        Dn = ctx.all_pairs_host(X, mode, refine=False)
        assert np.array_equal(Dn, raw), (np.argwhere(Dn != raw)[:8].tolist(), np.abs(Dn - raw).max())


# --------------------------------------------------------------------------------------------------
# --L2 on the tensor cores (tcgen05 Gram kernel + direct part + refinement)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cta_pairs", [False, True])
def test_gram_kernel_layouts(cta_pairs, monkeypatch):
    """cta_pairs: the default kernel on thread-block clusters of two (tcgen05 cta_group::2); False: one CTA per tile
    (FUF_GRAM_2CTA=0).
    Pins the shared-memory / instruction descriptors, the swizzle, the digit planes and the TMEM read-out of the
    integer Gram kernel: dense random profiles with column scales over six decades, (a) every column through the
    int8 path, (b) the default split with the largest-scale columns in float64.  A wrong layout gives O(1) errors;
    the arithmetic itself is exact up to 1.44e-9 of the scale statistic."""
    monkeypatch.setenv("FUF_GRAM_2CTA", "1" if cta_pairs else "0")
    n, N = 200, 300
    parent = np.full(n, n - 1, np.int32)
    parent[-1] = -1
    ctx = engine.TreeContext(parent, np.ones(n))
    rng = np.random.default_rng(0)
    P = (rng.random((n, N)) * (rng.random((n, N)) < 0.4)).astype(np.float32)
    P *= (10.0 ** rng.integers(-6, 1, (n, 1))).astype(np.float32)
    P[:, 7] = P[:, 3]                                                     # an identical pair
    P64 = P.astype(np.float64)
    exact = ((P64[:, :, None] - P64[:, None, :]) ** 2).sum(0)
    PT = torch.zeros((n, engine.padded_samples(N)), dtype=torch.float32, device="cuda")
    PT[:, :N] = torch.from_numpy(P)
    for direct in ("0", "16", None):
        if direct is None:
            monkeypatch.delenv("FUF_GRAM_DIRECT", raising=False)
        else:
            monkeypatch.setenv("FUF_GRAM_DIRECT", direct)
        D, h2, err = ctx.pairwise_gram(PT, N)
        D, h2, err = D.cpu().numpy(), h2.cpu().numpy()[:N], err.cpu().numpy()[:N]
        check_matrix(D)
        r = np.sqrt(err)[:, None] + np.sqrt(err)[None, :]
        qq = (P64 ** 2).sum(0)
        bound = (2 * np.sqrt(exact) * r + r ** 2 + 1.44e-9 * (h2[:, None] + h2[None, :])
                 + 1e-13 * (qq[:, None] + qq[None, :]))               # + float64 rounding of the sums themselves
        assert np.all(np.abs(D - exact) <= bound), direct                # the certified error model holds pairwise
        assert D[3, 7] == 0.0
        if direct == "0":
            assert np.abs(D - exact).max() / exact.max() < 1e-8 and h2.min() > 0


@pytest.mark.parametrize("cta_pairs", [False, True])
def test_gram_vs_oracle(cta_pairs, full_tree, monkeypatch):
    monkeypatch.setenv("FUF_GRAM_2CTA", "1" if cta_pairs else "0")
    inp, ctx, leaves = full_tree
    N = 640                                   # 5 tile rows, the last one 128 wide; above the Gram threshold
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=12)
    rng = np.random.default_rng(12)
    for base, copies, jitter in ((0, 4, 1e-3), (300, 3, 1e-5), (500, 3, 0.0)):
        for c in range(1, copies):
            X[base + c] = X[base] * (1 + jitter * rng.random(ctx.n))
            X[base + c] /= X[base + c].sum()
    Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, True), True)
    job = ctx.push_up_job(torch.from_numpy(X).cuda(), engine.MODE_L2)
    raw = ctx.pairwise_gram(job.PT, N, err_stat=job.err_stat)[0].cpu().numpy()
    D = ctx.distances(job, engine.MODE_L2)                      # N >= 512: routed to the Gram path
    assert 7 <= ctx.gram_refined <= 200
    D = D.cpu().numpy()
    check_matrix(D)
    assert rel_err(D, Dref) < REL
    ordinary = Dref > 1e-2 * np.median(Dref)
    assert (np.abs(raw - Dref)[ordinary] / Dref[ordinary]).max() < REL     # the tensor-core result itself
    assert rel_err(raw, Dref) > REL                                        # near-duplicates need the refinement
    assert raw[500, 501] == 0 and raw[501, 502] == 0                       # exact arithmetic: identical samples -> 0
    assert D[500, 501] == 0 and D[501, 502] == 0
    # correlated samples (mixtures): the floating-point Gram forms lose digits here, the integer one does not
    Xm = X.copy()
    for i in range(0, 200, 2):
        lam = 10 ** rng.uniform(-3, 0)
        Xm[i + 1] = (1 - lam) * Xm[i] + lam * Xm[i + 1]
    Drefm = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, Xm, True), True)
    Dm = ctx.distances(ctx.push_up_job(torch.from_numpy(Xm).cuda(), engine.MODE_L2), engine.MODE_L2).cpu().numpy()
    assert rel_err(Dm, Drefm) < REL and ctx.gram_refined < 400
    Dd = ctx.distances(job, engine.MODE_L2, gram=False).cpu().numpy()      # direct (a-b)^2 kernel
    assert rel_err(Dd, Dref) < REL
    Dh = ctx.all_pairs_host(X, engine.MODE_L2)                             # host path takes the Gram route too
    assert np.array_equal(Dh, D)
    Dh2 = ctx.all_pairs_host(X, engine.MODE_L2, gram=False)
    assert np.array_equal(Dh2, Dd)
    # compact tile rows on the all-gathered shard layout (the multi-GPU form), then assembled
    shard, G, T = 256, 3, 5
    PS = torch.zeros((G, ctx.n, shard), dtype=torch.float32, device="cuda")
    P64S = torch.zeros((G, ctx.n, shard), dtype=torch.float64, device="cuda")
    stat = torch.zeros(G * shard, dtype=torch.float64, device="cuda")
    Xd = torch.from_numpy(X).cuda()
    for g in range(G):
        w = min(shard, N - g * shard)
        ctx.push_up(Xd[g * shard:g * shard + w], engine.MODE_L2, PS[g], 0, job.center, None,
                    stat[g * shard:(g + 1) * shard])
        ctx.push_up_p64(Xd[g * shard:g * shard + w], engine.MODE_L2, P64S[g])
    blocks, rows = [], []
    for own in ([0, 3, 4], [1, 2]):
        blk, h2, err = ctx.pairwise_gram(PS, N, err_stat=stat, tile_rows=own, shard=shard, shard_stride=ctx.n * shard)
        ctx.refine(P64S, N, engine.MODE_L2, err, blk, tile_rows=own, shard=shard, shard_stride=ctx.n * shard,
                   lin_stat=h2, lin_kappa=engine.GRAM_KAPPA)
        blocks.append(blk)
        rows += own
    Ds = ctx.assemble(torch.cat(blocks, 0), rows, N).cpu().numpy()
    assert np.allclose(Ds, D, rtol=1e-12, atol=0) and rel_err(Ds, Dref) < REL
    from fununifrac_b200 import sharded
    one = sharded.ShardedAllPairs(sharded.CudaOps(ctx), N, engine.MODE_L2)
    assert one.use_gram and np.allclose(one.step(Xd).cpu().numpy(), D, rtol=1e-12, atol=0)
    # unweighted profiles: counts in the thousands on the upper levels
    Xu = X.copy()
    Xu[Xu > 0] = 1
    Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, Xu, True), True)
    D = ctx.distances(ctx.push_up_job(torch.from_numpy(Xu).cuda(), engine.MODE_L2), engine.MODE_L2).cpu().numpy()
    assert rel_err(D, Dref) < REL


# --------------------------------------------------------------------------------------------------
# (3) size-independent properties at a larger size + edge cases
# --------------------------------------------------------------------------------------------------
def test_properties_large(full_tree):
    inp, ctx, leaves = full_tree
    N = 1500
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=4, dtype=np.float32)
    Xd = torch.from_numpy(X).cuda()
    job = ctx.push_up_job(Xd, engine.MODE_L1)
    PT = ctx.push_up(Xd, engine.MODE_L1)
    D = ctx.distances(job, engine.MODE_L1).cpu().numpy()
    check_matrix(D)
    # homogeneity: scaling every profile by 2 is exact in binary floating point (the centre scales with it)
    D2 = ctx.distances(ctx.push_up_job(Xd * 2, engine.MODE_L1), engine.MODE_L1).cpu().numpy()
    assert np.array_equal(D2, 2 * D)
    # permutation equivariance, bit for bit, given the same centre (a pair's k-order does not depend on its
    # tile position)
    perm = np.random.default_rng(0).permutation(N)
    PTp = ctx.push_up(Xd[torch.from_numpy(perm).cuda()], engine.MODE_L1, center=job.center)
    Dp = ctx.pairwise(PTp, N, engine.MODE_L1).cpu().numpy()
    assert np.array_equal(Dp, ctx.pairwise(job.PT, N, engine.MODE_L1).cpu().numpy()[np.ix_(perm, perm)])
    # a different centre changes the fp32 operands but not the distances beyond the tolerance
    Dc = ctx.pairwise(ctx.push_up(Xd, engine.MODE_L1, center=ctx.push_up_center(Xd[100:], engine.MODE_L1, 7)), N,
                      engine.MODE_L1).cpu().numpy()
    assert rel_err(Dc, D) < REL
    # L1 on pushed vectors is a metric: triangle inequality on random triples
    rng = np.random.default_rng(1)
    i, j, k = rng.integers(0, N, (3, 20000))
    assert np.all(D[i, k] <= D[i, j] + D[j, k] + 1e-9)
    # a random sample of pairs against the oracle
    sub = np.sort(rng.choice(N, 64, replace=False))
    Pref = oracle_c.push_up(inp.parent, inp.edge_length, X[sub].astype(np.float64), False)
    assert rel_err(D[np.ix_(sub, sub)], oracle_c.pairwise(Pref, False)) < REL
    # mass conservation of the push-up: the root entry is the total mass
    assert np.allclose(PT[-1, :N].cpu().numpy(), X.astype(np.float64).sum(1), rtol=1e-6)


@pytest.mark.parametrize("N", [1, 2, 127, 128, 129])
def test_edge_sizes(N, full_tree):
    inp, ctx, leaves = full_tree
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=6)
    for l2 in (False, True):
        Pref = oracle_c.push_up(inp.parent, inp.edge_length, X, l2)
        D = ctx.all_pairs_host(X, engine.MODE_L2 if l2 else engine.MODE_L1)
        assert D.shape == (N, N) and rel_err(D, oracle_c.pairwise(Pref, l2)) < REL
        check_matrix(D)


def test_degenerate_trees():
    # single node (root only), chain (every node spans chunks), star; with zero-length edges
    rng = np.random.default_rng(0)
    chain = np.arange(1, 101, dtype=np.int32)
    chain[-1] = -1
    star = np.full(100, 99, dtype=np.int32)
    star[-1] = -1
    for parent in (np.array([-1], np.int32), chain, star):
        n = len(parent)
        length = rng.random(n) * (rng.random(n) > 0.2)
        ctx = engine.TreeContext(parent, length)
        X = rng.random((5, n))
        for l2 in (False, True):
            Pref = oracle_c.push_up(parent, length, X, l2)
            P64, P32, D64, D32 = run_all(ctx, X, l2)
            assert np.array_equal(P64, Pref)
            assert rel_err(D32, oracle_c.pairwise(Pref, l2)) < REL
    # empty input is a no-op, bad arguments are errors (never a crash)
    ctx = engine.TreeContext(chain, np.ones(100))
    assert ctx.all_pairs_host(np.zeros((0, 100)), engine.MODE_L1).shape == (0, 0)
    with pytest.raises(engine.FunUniFracError):
        engine.TreeContext(np.array([0, -1], np.int32), np.ones(2))       # parent[i] <= i
    with pytest.raises(ValueError):
        ctx.all_pairs_host(np.zeros((3, 99)), engine.MODE_L1)


def test_non_dfs_order_falls_back_to_exact_kernel():
    parent = np.array([3, 2, 4, 4, -1], np.int32)       # valid child<parent order, subtrees interleaved
    length = np.array([0.5, 0.25, 1.0, 2.0, 0.0])
    ctx = engine.TreeContext(parent, length)
    assert ctx.info().n_spanning == -1
    X = np.random.default_rng(0).random((4, 5))
    D = ctx.all_pairs_host(X, engine.MODE_L1)
    assert rel_err(D, oracle_c.pairwise(oracle_c.push_up(parent, length, X, False), False)) < REL


# --------------------------------------------------------------------------------------------------
# diffab, solver API, CLI
# --------------------------------------------------------------------------------------------------
def test_diffab_blocks_vs_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_mid.npz"))
    ctx = engine.TreeContext(z["parent"], z["length"])
    N, n = z["X"].shape
    Xd = torch.from_numpy(z["X"]).cuda()
    for tag, l2 in (("l1", False), ("l2", True)):
        mode = engine.MODE_L2 if l2 else engine.MODE_L1
        want = oracle_c.diffab_dense(z[f"{tag}.P"], l2)
        P64T = ctx.push_up_f64(Xd, mode)
        got = np.zeros((N, N, n))
        seen = 0
        for b, e, rows in engine.iter_diffab_blocks(ctx, P64T, N, mode, pairs_per_block=37):
            for p in range(b, e):
                i = int(np.searchsorted([engine.combinations_offset(r, N) for r in range(N)], p, side="right")) - 1
                j = p - engine.combinations_offset(i, N) + i + 1
                got[i, j] = rows[p - b]
                got[j, i] = -rows[p - b]
            seen += e - b
        assert seen == N * (N - 1) // 2
        assert np.array_equal(got, want)                 # float64 rows are bit-identical to P_i - P_j
        # the stream sharded over 3 "ranks": contiguous pair ranges that tile [0, total) and give the same rows
        whole = np.concatenate([rows.copy() for _, _, rows in engine.iter_diffab_blocks(ctx, P64T, N, mode, 64)])
        parts, edges = [], []
        for r in range(3):
            p0, p1 = engine.diffab_pair_range(N, 3, r)
            edges.append((p0, p1))
            for b, e, rows in engine.iter_diffab_blocks(ctx, P64T, N, mode, 50, pair_begin=p0, pair_end=p1):
                assert p0 <= b < e <= p1
                parts.append(rows.copy())
        assert edges[0][0] == 0 and edges[-1][1] == len(whole) and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert np.array_equal(np.concatenate(parts), whole)
        with pytest.raises(ValueError):
            next(engine.iter_diffab_blocks(ctx, P64T, N, mode, 50, pair_begin=5, pair_end=len(whole) + 1))
        PT = ctx.push_up(Xd, mode)
        out = torch.empty((N - 1, n), dtype=torch.float32, device="cuda")
        ctx.diffab_block(PT, N, mode, 0, N - 1, out)     # pairs (0, 1..N-1), fp32 production rows
        scale = np.abs(z[f"{tag}.P"]).max()
        assert np.abs(out.cpu().numpy() - want[0, 1:]).max() < 1e-6 * max(scale, scale ** 2)


def test_solver_api_matches_reference_semantics(toy, golden_dir):
    inp, ctx, z = toy
    solver = EarthMoverDistanceUniFracSolver()
    files = list(z["files"])
    for tag, l2 in (("l1_median", False), ("l2_median", True)):
        X, Pg, Dg, diffab_g = z[f"{tag}.X"], z[f"{tag}.P"], z[f"{tag}.D"], z[f"{tag}.diffab"]
        Ps = {}
        for f, x in zip(files, X):
            x0 = x.copy()
            Ps[f] = solver.push_up_L2(x, inp) if l2 else solver.push_up_L1(x, inp)
            assert np.array_equal(x, x0)                # the input vector is not modified
        assert np.array_equal(np.array([Ps[f] for f in files]), Pg)
        dists, none = solver.pairwise_computation(Ps, files, inp, False, l2)
        assert none is None and rel_err(dists, Dg) < REL and dists.dtype == np.float64
        dists, coo = solver.pairwise_computation(Ps, files, inp, True, l2)
        assert rel_err(dists, Dg) < REL
        assert np.array_equal(coo.todense(), diffab_g)      # float64 differences of the caller's own float64 vectors
        assert np.allclose(coo[2, 0, :].todense(), -coo[0, 2, :].todense())
        v = EarthMoverDistanceUniFracSolver(validate=True)
        dists, coo = v.pairwise_computation(Ps, files, inp, True, l2)
        assert rel_err(dists, Dg) < 1e-12 and np.array_equal(coo.todense(), diffab_g)
    # known diffab answers: tests/scripts/test_make_all_pw_fununifrac.py:78-88
    Ps = {f: solver.push_up_L1(x, inp) for f, x in zip(files, z["l1_median.X"])}
    _, coo = solver.pairwise_computation(Ps, files, inp, True, False)
    assert np.allclose(coo[0, 2, :].todense(), [-1 / 14, -1 / 21, -5 / 42, 5 / 42, 0, 5 / 42, 0], atol=1e-3)
    assert np.allclose(coo[0, 0, :].todense(), 0)
    # get_L1 & co
    from fununifrac_b200.objects import profile_vector as pv
    P, Q = Ps[files[0]], Ps[files[1]]
    assert np.isclose(pv.get_L1(P, Q), np.sum(np.abs(P - Q)), rtol=1e-14)
    assert np.isclose(pv.get_L2(P, Q), np.sum((P - Q) ** 2), rtol=1e-14)
    assert np.array_equal(pv.get_L1_diffab(P, Q), P - Q) and np.array_equal(pv.get_L2_diffab(P, Q), (P - Q) ** 2)


def test_push_up_side_effect_on_lint():
    from fununifrac_b200.objects.func_tree import FuncTreeEmduInput
    inp = FuncTreeEmduInput({0: 2, 1: 2}, {(0, 2): 0, (2, 0): 0, (1, 2): 1.0, (2, 1): 1.0}, [0, 1, 2],
                            {0: 'a', 1: 'b', 2: 'r'})
    out = EarthMoverDistanceUniFracSolver().push_up_L1(np.array([0.5, 0.5, 0.0]), inp)
    assert out.tolist() == [0.5 * 2.220446049250313e-16, 0.5, 1.0]
    assert inp.lint[(0, 2)] == 2.220446049250313e-16 and inp.lint[(2, 0)] == 0   # one direction only
    # the dicts are the caller's: a changed edge length must reach the device (the tree context is cached on the
    # input object and keyed on a checksum of the lengths)
    solver = EarthMoverDistanceUniFracSolver()
    assert solver.push_up_L1(np.array([0.25, 0.75, 0.0]), inp).tolist() == [0.25 * 2.220446049250313e-16, 0.75, 1.0]
    inp.lint[(1, 2)] = 3.0
    assert solver.push_up_L1(np.array([0.25, 0.75, 0.0]), inp).tolist() == [0.25 * 2.220446049250313e-16, 2.25, 1.0]
    assert solver.push_up_L2(np.array([0.0, 1.0, 0.0]), inp).tolist() == [0.0, np.sqrt(3.0), 1.0]


def test_legacy_solve_entry_points(data_dir, golden_dir):
    import json
    with open(os.path.join(golden_dir, "solve.json")) as f:
        g = json.load(f)
    solver = EarthMoverDistanceUniFracSolver()
    inp4 = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(
        os.path.join(data_dir, "small_edge_list_with_lengths_emdu.txt")))
    for weighted in (True, False):
        c = g[f"emdu4:weighted={weighted}"]
        Z, diffab = solver.solve(inp4, np.array(c["P"]), np.array(c["Q"]), weighted)
        assert Z == c["Z"] and {f"{a},{b}": v for (a, b), v in diffab.items()} == c["diffab"]
        Z, F, diffab = solver.solve_with_flow(inp4, np.array(c["P"]), np.array(c["Q"]), weighted)
        assert Z == c["Z_flow"] and {f"{a},{b}": float(v) for (a, b), v in F.items()} == c["F"]
        assert set(f"{a},{b}" for a, b in diffab) == set(c["diffab_flow"])
        for (a, b), v in diffab.items():
            assert np.isclose(v, c["diffab_flow"][f"{a},{b}"])
    inp7 = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(
        os.path.join(data_dir, "small_edge_list_with_lengths.txt")))
    for c in g["small7:dirichlet"]:
        Z, diffab = solver.solve(inp7, np.array(c["P"]), np.array(c["Q"]), True)
        assert Z == c["Z"] and {f"{a},{b}": v for (a, b), v in diffab.items()} == c["diffab"]


def test_cli_end_to_end(data_dir, golden_dir, tmp_path):
    """BASELINE config 1 through the CLI: same flags, same files, same contents as the reference run."""
    from fununifrac_b200 import compute_fununifrac
    from fununifrac_b200.utility import constant
    z = np.load(os.path.join(golden_dir, "toy_config1.npz"))
    edge = os.path.join(data_dir, "small_edge_list_with_lengths.txt")
    out = str(tmp_path / "out")
    base = ["-e", edge, "-fd", data_dir, "-o", out, "-i", "testing", "--force", "-b", "ko00001", "-a", "median_abund",
            "-t", "2"]
    compute_fununifrac.main(base)
    D = np.load(os.path.join(out, constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("testing")))
    assert D.dtype == np.float64 and rel_err(D, z["l1_median.D"]) < REL
    with open(os.path.join(out, constant.FUNUNIFRAC_OUT__MAIN_BASIS__FILE_NAME.format("testing"))) as f:
        basis = [os.path.basename(l.strip()) for l in f]
    assert basis == ["small_sim2_10_KOs_gather.csv", "small_sim3_10_KOs_gather.csv", "small_sim_10_KOs_gather.csv"]
    # --L2 --diffab --Ppushed
    compute_fununifrac.main(base + ["--L2", "--diffab", "--Ppushed"])
    D = np.load(os.path.join(out, constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("testing")))
    assert rel_err(D, z["l2_median.D"]) < REL
    coo = sparse_npz.load_npz(os.path.join(out, constant.FUNUNIFRAC_OUT__DIFFAB_FILE_NAME.format("testing")) + ".npz")
    assert coo.shape == (3, 3, 7) and np.allclose(coo.todense(), z["l2_median.diffab"], rtol=1e-13, atol=1e-16)
    assert np.allclose(coo[0, 2, :].todense(), [1 / 196, 1 / 441, 25 / 1764, 25 / 1764, 0, 25 / 1764, 0], atol=1e-3)
    with open(os.path.join(out, constant.FUNUNIFRAC_OUT__DIFFAB_NODES_FILE_NAME.format("testing"))) as f:
        assert [l.strip() for l in f] == ['d', 'e', 'b', 'f', 'g', 'c', 'ko00001']
    with open(os.path.join(out, constant.FUNUNIFRAC_OUT__PUSH_FILE_NAME.format("testing")), "rb") as f:
        pushed = pickle.load(f)
    files = sorted(glob.glob(os.path.join(data_dir, "*_gather.csv")))
    assert list(pushed) == files and np.array_equal(np.array([pushed[f] for f in files]), z["l2_median.P"])
    # --unweighted and the default abundance key
    compute_fununifrac.main(base + ["--unweighted"])
    D = np.load(os.path.join(out, constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("testing")))
    assert rel_err(D, z["l1_median_unweighted.D"]) < REL
    compute_fununifrac.main([a for a in base if a not in ("-a", "median_abund")] + ["--validate"])
    D = np.load(os.path.join(out, constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("testing")))
    assert rel_err(D, z["l1_default.D"], np.full((3, 3), 1e-9)) < 1e-12
    # validation errors of the driver
    with pytest.raises(ValueError):
        compute_fununifrac.main(base[:8] + ["-b", "not_a_brite", "-a", "median_abund"])
    with pytest.raises(Exception):
        compute_fununifrac.main(["-e", edge, "-fd", str(tmp_path), "-o", out])     # no matching files


def test_host_ranges_registered_in_pieces(full_tree, monkeypatch):
    """Some boxes page-lock a 20 GB result matrix only in pieces (fuf_host_register falls back to 1 GB registrations), and a
    copy-engine transfer must lie inside one registration: every pitched copy to or from such a range is cut at the piece
    boundaries.  Forced here with 64 KB pieces (FUF_REGISTER_STEP): rectangles whose rows straddle boundaries, then the
    whole host-buffer call — both modes, the Gram route with its completion segments — with X and D registered that way."""
    import ctypes
    import mmap
    inp, ctx, leaves = full_tree
    lib = engine.load_library()
    monkeypatch.setenv("FUF_REGISTER_STEP", str(64 << 10))

    def pieced(shape):
        nbytes = int(np.prod(shape)) * 8
        buf = mmap.mmap(-1, (nbytes + 4095) // 4096 * 4096)
        arr = np.frombuffer(buf, dtype=np.float64, count=int(np.prod(shape))).reshape(shape)
        pieces = ctypes.c_int32()
        engine._check(lib.fuf_host_register(arr.ctypes.data, nbytes, ctypes.byref(pieces)), "fuf_host_register")
        assert pieces.value == (nbytes + (64 << 10) - 1) // (64 << 10)
        return buf, arr, pieces

    def release(buf, arr, pieces):
        lib.fuf_host_unregister(arr.ctypes.data, arr.nbytes, pieces)

    N = 700                                   # a row of D is 5,600 bytes: 64 KB boundaries fall inside rows
    hb = pieced((N, N))
    Dh = hb[1]
    Dh[...] = -1.0
    src = torch.rand((N, N), dtype=torch.float64, device="cuda")
    target = ctx.host_device_pointer(Dh)
    rects = [((0, N), (0, N)), ((3, 400), (17, 650)), ((128, 256), (0, 128)), ((511, 700), (699, 700)), ((5, 6), (0, 700))]
    want = np.full((N, N), -1.0)
    for rows, cols in rects[1:]:
        ctx.copy_block_async(target, N, src, rows, cols)
        want[rows[0]:rows[1], cols[0]:cols[1]] = src[rows[0]:rows[1], cols[0]:cols[1]].cpu().numpy()
    torch.cuda.synchronize()
    assert np.array_equal(Dh, want)
    ctx.copy_block_async(target, N, src, *rects[0])
    torch.cuda.synchronize()
    assert np.array_equal(Dh, src.cpu().numpy())
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=29)
    hx = pieced(X.shape)
    hx[1][...] = X
    try:
        for l2 in (False, True):
            Dh[...] = np.nan
            ctx.all_pairs_host(hx[1], engine.MODE_L2 if l2 else engine.MODE_L1, Dh)
            ref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, l2), l2)
            assert rel_err(Dh, ref) < REL and np.array_equal(Dh, Dh.T)
    finally:
        release(*hx)
        release(*hb)


def test_integration_stub_runs(toy, full_tree):
    """INTEGRATION.md section B shows the ctypes binding a maintainer of the reference would add.  This test executes that
    very code block (library name replaced by the built library's path) on the toy tree and on the full synthetic tree,
    for both modes, against the oracle."""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(repo, "INTEGRATION.md")).read()
    code = "#" + text.split("```python\n# src/algorithms/emd_unifrac_b200.py", 1)[1].split("```", 1)[0]
    code = code.replace('ctypes.CDLL("libfununifrac.so")', f"ctypes.CDLL({engine.LIB_PATH!r})")
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    inp_toy, _, z = toy
    inp_full, ctx, leaves = full_tree
    for inp, X in ((inp_toy, z["l1_median.X"]), (inp_full, synth.synth_profiles_dense(leaves, ctx.n, 300, seed=23))):
        for l2 in (False, True):
            D = ns["all_pairs"](inp, list(X), l2)
            ref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, l2), l2)
            assert D.shape == ref.shape and rel_err(D, ref) < REL and np.array_equal(D, D.T)


def test_combined_table_cli(data_dir, golden_dir, tmp_path):
    """Second caller of the path (reference fununifrac/compute_fununifrac_combined.py): one KO x sample table.
    Expected matrices come from the reference itself, including its L1-push-up-with---L2 quirk."""
    from fununifrac_b200 import compute_fununifrac_combined as comb
    z = np.load(os.path.join(golden_dir, "combined.npz"))
    edge = os.path.join(data_dir, "small_edge_list_with_lengths.txt")
    table = os.path.join(data_dir, "small_combined_table.csv")
    out = str(tmp_path)
    comb.main(["-e", edge, "-f", table, "-o", out, "-b", "ko00001"])
    D = np.load(os.path.join(out, "fununifrac_out.npy"))
    assert D.dtype == np.float64 and rel_err(D, z["D_l1"]) < REL
    with open(os.path.join(out, "fununifrac_out.basis.npy")) as f:
        assert [l.strip() for l in f] == list(z["columns"])
    comb.main(["-e", edge, "-f", table, "-o", out, "-i", "t2", "--L2", "--diffab", "--unweighted", "--Ppushed"])
    assert rel_err(np.load(os.path.join(out, "t2.npy")), z["D_l2_quirk"]) < REL
    comb.main(["-e", edge, "-f", table, "-o", out, "-i", "t3", "--L2", "--validate"])
    assert rel_err(np.load(os.path.join(out, "t3.npy")), z["D_l2_quirk"]) < 1e-12


def test_many_uncertified_pairs_fall_back_to_float64(full_tree):
    """A data set dominated by near-duplicates: more pairs are flagged than per-pair refinement should handle
    (budget: 0.1 % of the pairs, at least 1,024), so the float64 tile kernel redoes the whole matrix."""
    inp, ctx, leaves = full_tree
    N = 520
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=13)
    rng = np.random.default_rng(13)
    for c in range(1, 70):                                   # 70 jittered copies of sample 0: 2,415 tiny distances
        X[c] = X[0] * (1 + 1e-4 * rng.random(ctx.n))
        X[c] /= X[c].sum()
    for l2 in (False, True):
        mode = engine.MODE_L2 if l2 else engine.MODE_L1
        Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, l2), l2)
        job = ctx.push_up_job(torch.from_numpy(X).cuda(), mode)
        D = ctx.distances(job, mode).cpu().numpy()
        assert ctx.last_flagged > 2000 and rel_err(D, Dref) < 1e-8       # float64 arithmetic on the float64 profiles
        Dh = ctx.all_pairs_host(X, mode)
        assert rel_err(Dh, Dref) < 1e-8


# ------------------------------------------------------------------------------------------------------
# diffab consumers on the GPU-resident store
# ------------------------------------------------------------------------------------------------------
def test_diffab_store_indexer_vs_reference_golden(toy):
    """The reference's test_diffab_indexer (tests/algorithms/test_emd_unifrac.py:142-184) against the store that keeps
    only the pushed matrix in HBM; values are those of the diffab tensor the unmodified reference produced."""
    from fununifrac_b200.utility import differential_abundance as da
    from test_diffab_consumers import check_indexer
    inp, ctx, z = toy
    files = [str(f) for f in z["files"]]
    for tag, l2 in (("l1_median", False), ("l2_median", True), ("l2_median_unweighted", True)):
        Ps = {f: p for f, p in zip(files, z[f"{tag}.P"])}
        store = da.PushedDiffabStore.from_pushed(Ps, files, inp, l2)
        golden = z[f"{tag}.diffab"]
        assert store.shape == golden.shape
        assert np.array_equal(store.todense(), golden)               # float64 differences of the same pushed vectors
        check_indexer(da.DiffabArrayIndexer(store, inp.basis, files, inp.idx_to_node), golden, files, atol=0)
        assert np.array_equal(store[0, 2], golden[0, 2]) and np.array_equal(store[2, 0, :], golden[2, 0, :])
        assert store[1, 2, 3] == golden[1, 2, 3] and np.array_equal(store[-1], golden[-1])
        assert np.array_equal(store[0:2, [2, 0], 1:5], golden[0:2][:, [2, 0]][:, :, 1:5])
        with pytest.raises(IndexError):
            store[3, 0]
        df = da.convert_diffab_array_to_df(store, [inp.idx_to_node[i] for i in inp.basis], files)
        assert df[files[0]][files[2]]["d"] == golden[0, 2, 0]


@pytest.mark.parametrize("l2", [False, True])
def test_diffab_gather_vs_oracle_blocks(l2, full_tree):
    """Ragged index lists on the 30,002-node tree: every block equals the dense definition built from the oracle's
    pushed vectors (antisymmetric; L2 squared with the sign of j - i)."""
    inp, ctx, leaves = full_tree
    N = 70
    mode = engine.MODE_L2 if l2 else engine.MODE_L1
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=11)
    P = oracle_c.push_up(inp.parent, inp.edge_length, X, l2)
    P64T = ctx.push_up_f64(torch.from_numpy(X).cuda(), mode)
    assert np.array_equal(P64T[:, :N].cpu().numpy().T, P)
    rng = np.random.default_rng(4)
    for na, nb, nk in ((1, 1, None), (3, 70, None), (33, 65, 100), (70, 2, 1), (5, 40, 4097)):
        ri, rj = rng.integers(0, N, na), rng.integers(-N, N, nb)
        nodes = None if nk is None else rng.integers(0, ctx.n, nk)
        got = ctx.diffab_gather(P64T, N, mode, ri, rj, nodes).cpu().numpy()
        d = P[ri][:, None, :] - P[rj % N][None, :, :]
        if l2:
            d = np.sign((rj % N)[None, :] - ri[:, None])[:, :, None] * d * d
        want = d if nodes is None else d[:, :, nodes]
        assert got.shape == want.shape and np.array_equal(got, want)
    out32 = torch.empty((2, 3, ctx.n), dtype=torch.float32, device="cuda")
    ctx.diffab_gather(P64T, N, mode, [0, 1], [2, 3, 4], None, out=out32)
    d = P[[0, 1]][:, None, :] - P[[2, 3, 4]][None, :, :]
    assert np.array_equal(out32.cpu().numpy(), (d * d if l2 else d).astype(np.float32))
    with pytest.raises(IndexError):
        ctx.diffab_gather(P64T, N, mode, [N], [0])
    assert ctx.diffab_gather(P64T, N, mode, [], [0, 1]).shape == (0, 2, ctx.n)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("l2", [False, True])
def test_cli_under_torchrun_two_gpus(l2, tmp_path):
    """The drop-in CLI launched with torchrun: every rank ingests its block of files, tile rows are sharded, the ranks
    store into one shared host matrix, rank 0 writes the reference's output files — same matrix as the 1-GPU run."""
    import subprocess
    import sys
    from fununifrac_b200 import compute_fununifrac
    from fununifrac_b200.utility import constant
    tree = synth.synth_ko_tree(str(tmp_path / "tree.txt"), seed=3, widths=(1, 1, 4, 12, 40, 160, 900))
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(tree), "ko00001")
    n = len(inp.basis)
    leaves = synth.leaves_of(inp.parent)
    names = [inp.idx_to_node[k] for k in range(n)]
    prof = tmp_path / "profiles"
    prof.mkdir()
    N = 600 if l2 else 300                 # --L2 from 512 samples on: the tensor-core Gram route
    rng = np.random.default_rng(8)
    for i in range(N):
        indptr, indices, values = synth.synth_profile_rows(leaves, n, [i], seed=21)
        synth.write_gather_csv(str(prof / f"s{i:04d}_gather.csv"), [names[j] for j in indices],
                               rng.geometric(0.2, len(indices)))
    base = ["-e", tree, "-fd", str(prof), "-b", "ko00001", "-a", "median_abund", "-t", "4"] + (["--L2"] if l2 else [])
    compute_fununifrac.main(base + ["-o", str(tmp_path / "one"), "-i", "x"])
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=repo + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", "-m", "fununifrac_b200.compute_fununifrac"] + base + \
          ["-o", str(tmp_path / "two"), "-i", "x"]
    res = subprocess.run(cmd, env=env, cwd=repo, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    name = constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("x")
    D1, D2 = np.load(tmp_path / "one" / name), np.load(tmp_path / "two" / name)
    assert D2.shape == (N, N) and np.array_equal(D2, D2.T) and not np.diag(D2).any()
    assert rel_err(D2, D1) < 1e-9          # same kernels; only the order of a few float64 additions differs
    P = oracle_c.push_up(inp.parent, inp.edge_length,
                         make_emd_input.native_profiles_to_matrix(sorted(glob.glob(str(prof / "*_gather.csv"))), inp,
                                                                  "median_abund", False, 4)[1], l2)
    assert rel_err(D2, oracle_c.pairwise(P, l2)) < REL
    bn = constant.FUNUNIFRAC_OUT__MAIN_BASIS__FILE_NAME.format("x")
    assert open(tmp_path / "one" / bn).read() == open(tmp_path / "two" / bn).read()


@pytest.mark.parametrize("ranks,N,mode", [(2, 900, "l1"), (3, 700, "l1"), (4, 400, "l2"), (2, 640, "l2"), (3, 1100, "l2"),
                                          (2, 2100, "l1")] +       # (banded upload of the own shard)
                         ([(3, 1700, "l1")] if os.environ.get("FUF_TEST_BAND_PULLS") == "1" else []))
def test_sharded_job_ranks_sharing_one_gpu(ranks, N, mode):
    """The multi-GPU job with several ranks on ONE device (collectives over gloo, staged through the host): everything
    CUDA-side of the peer form is real — IPC mappings between the processes, copy-engine pulls, arrival flags polled by
    the running kernel, tiles stored in place into rank 0's matrix and into the shared host matrix, the float64 list
    pass — so the form is covered on the driver's single-GPU box too."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # FUF_SM_LIMIT: on ONE device the "peer" pulls are same-device copies, which the driver may run as SM kernels; with a
    # persistent CTA spinning for its shard on every SM such a copy could never start (measured: N = 2,100 deadlocks into
    # the kernels' bounded-wait trap).  Between real GPUs the pulls are copy-engine transfers and need no SM.
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0], FUF_SM_LIMIT="140")
    if N == 1700 and os.environ.get("FUF_TEST_BAND_PULLS") == "1":
        # the experimental band-wise pulls of the host-buffer form (off by default; green here and on two real B200s in
        # round 2 — opt-in, so that the default suite exercises exactly what ships)
        env["FUF_BAND_PULLS"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr",
           "127.0.0.1", "--master-port", "29543", os.path.join(repo, "tests", "multi_gpu_worker.py"), str(N), mode]
    res = subprocess.run(cmd, cwd=repo, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "MULTI_GPU_OK" in res.stdout, (res.stdout[-2000:], res.stderr[-4000:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("N,mode", [(1500, "l1"), (300, "l2"), (700, "l2"), (130, "l1")])
def test_sharded_job_on_real_gpus(N, mode):
    """sharded.ShardedAllPairs on every visible GPU (tests/multi_gpu_worker.py under torchrun): peer form for weighted L1
    and the direct --L2 kernel, collective form for the Gram path (N >= 512), device-resident and host-buffer results,
    near-duplicate and identical pairs, against the oracle."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = torch.cuda.device_count()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(g), "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(repo, "tests", "multi_gpu_worker.py"), str(N), mode]
    res = subprocess.run(cmd, cwd=repo, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "MULTI_GPU_OK" in res.stdout, (res.stdout[-2000:], res.stderr[-4000:])


# ------------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes (configs 3 and 4) through size-independent properties
# ------------------------------------------------------------------------------------------------------
def _full_size_profiles(leaves, n, N, seed):
    """Uniform abundances on a random 20 % of the leaves, normalised, generated on the device (bench.py's generator)."""
    import bench
    return bench.device_profiles(leaves, n, N, seed=seed)


def test_full_size_config3_l1_10k(full_tree):
    """10,000 samples x 30,002 nodes, weighted L1 (the headline workload): structure of the whole matrix, planted
    identical / near-duplicate samples, and a random 320-sample block against the oracle."""
    inp, ctx, leaves = full_tree
    N = 10000
    X = _full_size_profiles(leaves, ctx.n, N, seed=31)
    X[7777] = X[12]                                                   # identical pair, far apart
    X[4242] = X[13] * (1 + 1e-7 * torch.rand(ctx.n, device="cuda", dtype=torch.float64))   # near-duplicate: must be refined
    job = ctx.push_up_job(X, engine.MODE_L1)
    D = ctx.distances(job, engine.MODE_L1)
    assert ctx.refine_count() >= 2
    assert torch.equal(D, D.T) and not bool(torch.diagonal(D).any()) and bool(torch.isfinite(D).all())
    assert float(D[12, 7777]) == 0.0
    off = D + torch.eye(N, device="cuda", dtype=torch.float64)
    off[12, 7777] = off[7777, 12] = 1.0
    assert float(off.min()) > 0.0                                     # every other pair is strictly apart
    # mass conservation: the root entry of every pushed profile is its total mass (1 up to rounding)
    assert torch.allclose(job.P64T[-1, :N], X.sum(1), rtol=1e-14)
    rng = np.random.default_rng(2)
    sub = np.sort(np.concatenate([rng.choice(N, 316, replace=False), [12, 13, 4242, 7777]]))
    sub = np.unique(sub)
    Pref = oracle_c.push_up(inp.parent, inp.edge_length, X[torch.from_numpy(sub).cuda()].cpu().numpy(), False)
    Dref = oracle_c.pairwise(Pref, False)
    got = D[torch.from_numpy(sub).cuda()][:, torch.from_numpy(sub).cuda()].cpu().numpy()
    i12, i7777 = int(np.searchsorted(sub, 12)), int(np.searchsorted(sub, 7777))
    scale = np.full_like(Dref, 1e-300)
    scale[i12, i7777] = scale[i7777, i12] = 1.0                       # the identical pair: 0 against the oracle's 0
    assert rel_err(got, Dref, scale) < REL                            # includes the near-duplicate pair (13, 4242)
    # triangle inequality on random triples of the full matrix
    i, j, k = (torch.from_numpy(rng.integers(0, N, 200000)).cuda() for _ in range(3))
    assert bool((D[i, k] <= D[i, j] + D[j, k] + 1e-9).all())


def test_full_size_config4_l2_50k(full_tree):
    """50,000 samples, --L2 on the tensor cores: every one of the 1.25e9 distances enters a row-sum identity that
    float64 can evaluate in O(N n):  sum_j ||a_i - a_j||^2 = N q_i + sum_j q_j - 2 a_i . (sum_j a_j)."""
    inp, ctx, leaves = full_tree
    N = 50000
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 90 * (1 << 30):
        pytest.skip("needs ~80 GB of free device memory")
    X = _full_size_profiles(leaves, ctx.n, N, seed=32)
    X[31000] = X[5]
    # a block for the oracle, rows and columns far apart in the matrix: its un-pushed profiles go to the host now
    rng = np.random.default_rng(3)
    sub = np.unique(np.concatenate([rng.choice(N, 256, replace=False), [5, 31000]]))
    idx = torch.from_numpy(sub).cuda()
    Xsub = X[idx].cpu().numpy()
    job = ctx.push_up_job(X, engine.MODE_L2)
    del X
    D = ctx.distances(job, engine.MODE_L2)
    assert float(D[5, 31000]) == 0.0 and float(D[31000, 5]) == 0.0
    assert not bool(torch.diagonal(D).any())
    rows = D.sum(1)
    A = job.P64T[:, :N]
    q = (A * A).sum(0)
    s = A.sum(1)
    want = N * q + q.sum() - 2.0 * (s @ A)
    assert float(((rows - want).abs() / want).max()) < 1e-7
    cols = D.sum(0)
    assert torch.equal(rows, cols) or float(((rows - cols).abs() / rows).max()) < 1e-14     # symmetric (sum order)
    # the block against the oracle's own push-up + pairwise (nothing of this repository in the expected values)
    Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, Xsub, True), True)
    got = D[idx][:, idx].cpu().numpy()
    i5, i31 = int(np.searchsorted(sub, 5)), int(np.searchsorted(sub, 31000))
    scale = np.full_like(Dref, 1e-300)
    scale[i5, i31] = scale[i31, i5] = 1.0
    assert rel_err(got, Dref, scale) < REL
    del D, job, A
    torch.cuda.empty_cache()
