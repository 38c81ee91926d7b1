"""CPU checks of the device algorithms: the library's own push-up plan + numpy models of the kernels
(tests/kernel_models.py) against the oracle; plus the ABI surface and the host-side containers."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from fununifrac_b200 import engine, sparse_npz, synth
from fununifrac_b200.algorithms.flow import flow_matching
from fununifrac_b200.factory import make_emd_input, make_tree
from oracle import oracle_c, oracle_np

import kernel_models as km

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = engine.load_library()
    hdr = open(os.path.join(REPO, "include", "fununifrac.h")).read()
    syms = sorted(set(re.findall(r"\b(fuf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(L, s), s
    assert L.fuf_version() == engine.ABI_VERSION == int(re.search(r"#define FUF_ABI_VERSION (\d+)", hdr).group(1))
    assert L.fuf_padded_samples(1) == 128 and L.fuf_padded_samples(129) == 256


@pytest.mark.parametrize("n_tiles,sms", [(3160, 148), (405, 148), (36, 148), (1, 148), (148, 148), (149, 148), (295, 148),
                                         (820, 140), (7, 4)])
def test_pair_work_plan_covers_every_stage_once(n_tiles, sms):
    """The work plan of the all-pairs tile kernel (whole tiles for the full waves + a stream-K tail): every (tile, stage)
    exactly once, span boundaries on fp32 flush boundaries, shared tiles' slots in ascending stage order, and no CTA
    more than a few flush units above the mean load."""
    import ctypes
    L = engine.load_library()
    stages, flush = 1876, 16
    for stream_k in (1, 0):
        nr, grid = ctypes.c_int32(), ctypes.c_int32()
        assert L.fuf_debug_pair_plan(n_tiles, sms, stages, flush, stream_k, None, 0, ctypes.byref(nr), ctypes.byref(grid)) == 0
        rows = np.zeros((nr.value, 5), np.int32)
        assert L.fuf_debug_pair_plan(n_tiles, sms, stages, flush, stream_k, rows.ctypes.data, nr.value, ctypes.byref(nr),
                                     ctypes.byref(grid)) == 0
        assert 1 <= grid.value <= sms and rows[:, 0].max() == grid.value - 1
        cover = np.zeros((n_tiles, stages), np.int32)
        load = np.zeros(grid.value, np.int64)
        slots = {}
        for c, t, s0, s1, slot in rows:
            assert 0 <= s0 < s1 <= stages and s0 % flush == 0 and (s1 % flush == 0 or s1 == stages)
            cover[t, s0:s1] += 1
            load[c] += s1 - s0
            whole = s0 == 0 and s1 == stages
            assert (slot < 0) == whole
            if slot >= 0:
                slots.setdefault(t, []).append((slot, s0))
        assert (cover == 1).all()
        for t, lst in slots.items():            # consecutive slots, ascending stages
            assert [a for a, _ in lst] == list(range(lst[0][0], lst[0][0] + len(lst))) and [b for _, b in lst] == sorted(b for _, b in lst)
        assert sorted(a for lst in slots.values() for a, _ in lst) == list(range(sum(len(v) for v in slots.values())))
        if stream_k:
            # no CTA more than a minimal span (4 flush units) + rounding above a perfect deal over all SMs
            assert load.max() <= n_tiles * stages / sms + 5 * flush, (load.max(), n_tiles * stages / sms)
        else:
            assert not slots


@pytest.mark.parametrize("T,n_seg", [(391, 32), (391, 0), (64, 32), (5, 3), (1, 1), (129, 16), (40, 64)])
def test_gram_item_order_and_segments(T, n_seg):
    """Items of a full-matrix launch of the CTA-pair Gram kernel: every (row tile, column-tile pair) of the upper triangle
    exactly once; rasterised in strips that are a whole number of completion segments; a segment's counter target is its
    item count x 16; any 74 consecutive items (the CTA pairs in flight) touch few distinct panels."""
    import ctypes
    L = engine.load_library()
    nr, ns = ctypes.c_int32(), ctypes.c_int32()
    assert L.fuf_debug_gram_items(T, n_seg, None, 0, ctypes.byref(nr), None, None, 0, ctypes.byref(ns)) == 0
    rows = np.zeros((nr.value, 3), np.int32)
    row_end, expected = np.zeros(max(1, ns.value), np.int32), np.zeros(max(1, ns.value), np.int32)
    assert L.fuf_debug_gram_items(T, n_seg, rows.ctypes.data, nr.value, ctypes.byref(nr), row_end.ctypes.data,
                                  expected.ctypes.data, len(row_end), ctypes.byref(ns)) == 0
    want = {(t, b) for t in range(T) for b in range(t // 2, (T + 1) // 2)}
    assert len(rows) == len(want) and {(int(t), int(b)) for t, b, _ in rows} == want
    if n_seg == 0:
        assert ns.value == 0 and (rows[:, 2] == -1).all()
    else:
        assert 1 <= ns.value <= min(n_seg, T) and row_end[ns.value - 1] == T and (np.diff(row_end[:ns.value]) > 0).all()
        per = int(row_end[0])
        assert (rows[:, 2] == rows[:, 0] // per).all()
        counts = np.bincount(rows[:, 2], minlength=ns.value)
        assert (expected[:ns.value] == counts * engine.GRAM_SEG_UNITS).all()
        first = [int(np.argmax(rows[:, 2] == s)) for s in range(ns.value)]
        last = [len(rows) - 1 - int(np.argmax(rows[::-1, 2] == s)) for s in range(ns.value)]
        assert last == sorted(last)                                  # segments complete in order
        strip = per * -(-12 // per)
        assert all(first[s] > last[s - strip // per] for s in range(strip // per, ns.value))   # ... strip by strip
    if T >= 128:                                                     # locality of the 74 item pairs in flight
        worst = 0
        for at in range(0, len(rows) - 74, 997):
            win = rows[at:at + 74]
            worst = max(worst, len(set(win[:, 0])) + 2 * len(set(win[:, 1])))
        assert worst <= 64, worst                                    # (a window that straddles two strips; row-major: 149)


def test_integration_stub_compiles_and_binds():
    """The ctypes block of INTEGRATION.md (executed against the oracle by the GPU test test_integration_stub_runs) is valid
    Python and every entry point it binds is exported — checked here without a GPU."""
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    code = "#" + text.split("```python\n# src/algorithms/emd_unifrac_b200.py", 1)[1].split("```", 1)[0]
    code = code.replace('ctypes.CDLL("libfununifrac.so")', f"ctypes.CDLL({engine.LIB_PATH!r})")
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    assert callable(ns["all_pairs"])
    for name in re.findall(r"_L\.(fuf_[a-z0-9_]+)", code):
        assert hasattr(ns["_L"], name), name


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.FunUniFracError):
        engine.TreeContext(np.array([1, -1], np.int32), np.array([1.0, 0.0]))
    # and the raw ABI reports ENODEV rather than computing anything on the CPU
    L = engine.load_library()
    h = ctypes.c_void_p()
    p = np.array([1, -1], np.int32)
    l = np.array([1.0, 0.0])
    rc = L.fuf_tree_create(ctypes.byref(h), -1, p.ctypes.data, l.ctypes.data, 2)
    assert rc == -4 and b"no CPU fallback" in L.fuf_last_error()


def _random_postorder_tree(rng, n, max_children=6):
    """Random tree, relabelled into DFS postorder (root last)."""
    par = np.full(n, -1)
    for v in range(1, n):
        par[v] = rng.integers(max(0, v - 40), v) if rng.random() < 0.7 else rng.integers(0, v)
    kids = [[] for _ in range(n)]
    for v in range(1, n):
        kids[par[v]].append(v)
    order, stack = [], [(0, 0)]
    while stack:
        v, i = stack.pop()
        if i < len(kids[v]):
            stack.append((v, i + 1))
            stack.append((kids[v][i], 0))
        else:
            order.append(v)
    new = np.empty(n, dtype=np.int64)
    new[order] = np.arange(n)
    parent = np.full(n, -1, dtype=np.int32)
    for v in range(1, n):
        parent[new[v]] = new[par[v]]
    return parent


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 0), (7, 1), (31, 2), (32, 3), (33, 4), (64, 5), (257, 6), (1000, 7),
                                    (4097, 8)])
def test_pushup_plan_and_model_random_trees(n, seed):
    rng = np.random.default_rng(seed)
    parent = _random_postorder_tree(rng, n)
    length = rng.random(n) * (rng.random(n) > 0.1)      # some exact zeros
    X = rng.random((5, n)) * (rng.random((5, n)) > 0.5)
    plan = km.tree_plan(parent)
    assert plan["contiguous"]
    for l2 in (False, True):
        want = oracle_c.push_up(parent, length, X, l2)
        got = km.pushup_model(parent, length, X, l2, plan)
        assert np.allclose(got, want, rtol=3e-7, atol=1e-30)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 0), (7, 1), (32, 3), (33, 4), (257, 6), (1000, 7), (4097, 8)])
def test_tma_pushup_plan_words_random_trees(n, seed):
    """The plan the TMA push-up kernel runs on (leaf-like / skip / parked-prefix / checkpoint bits per node): a numpy
    replay of the kernel's arithmetic from those words reproduces the oracle within fp32 rounding, centred or not, and
    its statistic is the rounding error it actually made."""
    rng = np.random.default_rng(seed)
    parent = _random_postorder_tree(rng, n)
    length = rng.random(n) * (rng.random(n) > 0.1)
    X = rng.random((4, n)) * (rng.random((4, n)) > 0.5)
    meta = km.push_meta(parent)
    assert meta is not None and len(meta) % 32 == 0 and np.all(meta[n:] == 2)
    has_child = np.zeros(n, bool)
    has_child[parent[parent >= 0]] = True
    assert np.array_equal((meta[:n] & 1).astype(bool) & ~(meta[:n] & 2).astype(bool),
                          ~has_child & ~(meta[:n] & 2).astype(bool))        # leaf-like = no children
    for l2 in (False, True):
        want = oracle_c.push_up(parent, length, X, l2)
        got, err = km.pushup_tma_model(parent, length, X, l2)
        assert np.allclose(got, want, rtol=3e-7, atol=1e-30)
        center = np.where(has_child, want[:2].mean(0), 0.0)
        got_c, err_c = km.pushup_tma_model(parent, length, X, l2, center=center)
        resid = (want - center) - got_c.astype(np.float64)
        stat = (resid ** 2).sum(1) if l2 else np.abs(resid).sum(1)
        assert np.allclose(got_c, want - center, rtol=3e-7, atol=1e-7 * np.abs(want).max())
        assert np.allclose(err_c, stat, rtol=1e-6, atol=1e-12 * stat.max() + 1e-300)


def test_tma_pushup_plan_unusable_trees():
    assert km.push_meta(np.array([3, 2, 4, 4, -1], np.int32)) is None         # not a DFS postorder


def test_pushup_model_chain_and_star():
    # a path graph (depth n-1: every node spans) and a star (root spans everything)
    n = 300
    chain = np.arange(1, n + 1, dtype=np.int32)
    chain[-1] = -1
    star = np.full(n, n - 1, dtype=np.int32)
    star[-1] = -1
    rng = np.random.default_rng(0)
    X = rng.random((3, n))
    length = rng.random(n)
    for parent in (chain, star):
        want = oracle_c.push_up(parent, length, X, False)
        assert np.allclose(km.pushup_model(parent, length, X, False), want, rtol=3e-7)


def test_non_dfs_order_is_flagged_not_contiguous():
    # child < parent holds, but subtrees are interleaved: 0->3, 1->2, 2->4, 3->4
    parent = np.array([3, 2, 4, 4, -1], np.int32)
    assert not km.tree_plan(parent)["contiguous"]


def test_pushup_model_on_golden_trees(golden_dir, real_trees):
    z = np.load(os.path.join(golden_dir, "synth_mid.npz"))
    for tag, l2 in (("l1", False), ("l2", True)):
        got = km.pushup_model(z["parent"], z["length"], z["X"], l2)
        assert np.allclose(got, z[f"{tag}.P"], rtol=2e-7, atol=1e-30)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(
        real_trees["kegg_tree_scale_10_k_5_deterministic.txt"]), "ko00001")
    plan = km.tree_plan(inp.parent)
    assert plan["contiguous"] and 0 < len(plan["span"]) < 2000
    leaves = synth.leaves_of(inp.parent)
    X = synth.synth_profiles_dense(leaves, len(inp.parent), 3, seed=5)
    want = oracle_c.push_up(inp.parent, inp.edge_length, X, False)
    assert np.allclose(km.pushup_model(inp.parent, inp.edge_length, X, False, plan), want, rtol=2e-7, atol=1e-30)
    # float32 input (the benchmark's resident format) stays within fp32 rounding of the float64 reference
    got32 = km.pushup_model(inp.parent, inp.edge_length, X, False, plan, x_dtype=np.float32)
    assert np.allclose(got32, want, rtol=3e-7, atol=1e-30)


def test_pairwise_numerical_scheme_meets_1e6(golden_dir):
    """fp32 operands + 256-entry fp32 partial sums folded into double stay within 1e-6 relative of the
    reference's float64 distances on the golden set (incl. near-duplicates relative to their scale)."""
    z = np.load(os.path.join(golden_dir, "synth_mid.npz"))
    for tag, l2 in (("l1", False), ("l2", True)):
        P32 = km.pushup_model(z["parent"], z["length"], z["X"], l2)
        D = km.pairwise_model(P32, l2)
        ref = z[f"{tag}.D"]
        iu = np.triu_indices(len(ref), 1)
        rel = np.abs(D[iu] - ref[iu]) / np.maximum(ref[iu], 1e-300)
        ordinary = ref[iu] > 1e-2 * np.median(ref[iu])
        assert rel[ordinary].max() < 1e-6
        assert D[1, 21] == 0.0 and ref[1, 21] == 0.0          # exact duplicates give exactly 0
        # near-duplicate pair (0, 20): error relative to the operand scale
        scale = np.abs(z[f"{tag}.P"][0]).sum() if not l2 else (z[f"{tag}.P"][0] ** 2).sum()
        assert abs(D[0, 20] - ref[0, 20]) < 1e-6 * scale


def test_flow_matching_matches_reference_goldens(golden_dir):
    with open(os.path.join(golden_dir, "solve.json")) as f:
        g = json.load(f)
    with open(os.path.join(golden_dir, "trees.json")) as f:
        trees = json.load(f)
    t4 = trees["small_edge_list_with_lengths_emdu.txt:nobrite"]
    for weighted in (True, False):
        c = g[f"emdu4:weighted={weighted}"]
        Z, F, reached = flow_matching(np.array(t4["parent"]), np.array(t4["length"]), np.array(c["P"]), np.array(c["Q"]))
        assert Z == c["Z_flow"]
        assert {f"{a},{b}": float(v) for (a, b), v in F.items()} == c["F"]
        assert {f"{i},{t4['parent'][i]}" for i in range(3) if reached[i]} == set(c["diffab_flow"])
    t7 = trees["small_edge_list_with_lengths.txt:nobrite"]
    for c in g["small7:dirichlet"]:
        Z, F, reached = flow_matching(np.array(t7["parent"]), np.array(t7["length"]), np.array(c["P"]), np.array(c["Q"]))
        assert np.isclose(Z, c["Z_flow"], rtol=1e-13)
        assert set(f"{a},{b}" for a, b in F) == set(c["F"])
        for (a, b), v in F.items():
            assert np.isclose(v, c["F"][f"{a},{b}"], rtol=1e-12, atol=1e-15)


def test_coo_container_and_npz_roundtrip(tmp_path, golden_dir):
    z = np.load(os.path.join(golden_dir, "toy_config1.npz"))
    dense = z["l1_median.diffab"]                      # reference: COO - COO.transpose((1,0,2)), densified
    N, _, n = dense.shape
    up = np.zeros_like(dense)
    for i in range(N):
        for j in range(i + 1, N):
            up[i, j] = dense[i, j]
    coo_up = sparse_npz.COO(np.array(np.nonzero(up)), up[np.nonzero(up)], up.shape)
    full = coo_up - coo_up.transpose((1, 0, 2))
    assert np.array_equal(full.todense(), dense)
    assert np.array_equal(full[0, 2, :].todense(), dense[0, 2]) and np.array_equal(full[2, 0, :].todense(), -dense[0, 2])
    assert np.all(np.diff(np.ravel_multi_index(full.coords, full.shape)) > 0)   # sorted, no duplicates
    path = tmp_path / "x.diffab"
    sparse_npz.save_npz(str(path), full)
    assert os.path.exists(str(path) + ".npz")           # numpy appends .npz like sparse.save_npz does
    with np.load(str(path) + ".npz") as fp:
        assert sorted(fp.files) == ["coords", "data", "fill_value", "shape"]
    back = sparse_npz.load_npz(str(path) + ".npz")
    assert np.array_equal(back.todense(), dense) and back.shape == (N, N, n)
    # the numpy oracle's COO agrees
    coords, data = oracle_np.diffab_coo(list(z["l1_median.P"]), False)
    assert np.array_equal(coords, full.coords) and np.array_equal(data, full.data)


def test_pair_row_lookup():
    from fununifrac_b200.algorithms.emd_unifrac import _pair_rows
    import itertools
    for N in (2, 3, 5, 64, 501):
        pairs = np.array(list(itertools.combinations(range(N), 2)))
        p = np.arange(len(pairs))
        i = _pair_rows(p, N)
        assert np.array_equal(i, pairs[:, 0])
        assert np.array_equal(p - (i * N - i * (i + 1) // 2) + i + 1, pairs[:, 1])
        assert all(engine.combinations_offset(r, N) == np.searchsorted(pairs[:, 0], r) for r in range(N - 1))
