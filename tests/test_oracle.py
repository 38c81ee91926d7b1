"""Pins the CPU oracle (oracle/oracle.c, oracle/oracle_np.py) against outputs of the unmodified
reference (tests/golden/make_golden.py).  Bit-exact: push-up replays the reference loop, distances
restate numpy's pairwise summation."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle_c, oracle_np


def sha(x):
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def toy(golden_dir):
    with open(os.path.join(golden_dir, "trees.json")) as f:
        t = json.load(f)["small_edge_list_with_lengths.txt"]
    return np.array(t["parent"], np.int32), np.array(t["length"]), np.load(os.path.join(golden_dir, "toy_config1.npz"))


def test_np_sum_restatement():
    rng = np.random.default_rng(0)
    for n in [0, 1, 7, 8, 9, 127, 128, 129, 1000, 25961, 30002, 100003]:
        a = rng.random(n) * 10.0 ** rng.integers(-8, 3, size=n)
        assert oracle_c.np_sum(a) == np.sum(a), n


@pytest.mark.parametrize("tag,l2", [("l1_median", False), ("l2_median", True), ("l1_median_unweighted", False),
                                    ("l2_median_unweighted", True), ("l1_default", False)])
def test_toy_config1(tag, l2, toy):
    parent, length, z = toy
    X, P, D, diffab = z[f"{tag}.X"], z[f"{tag}.P"], z[f"{tag}.D"], z[f"{tag}.diffab"]
    for impl in (oracle_c, ):
        Pc = impl.push_up(parent, length, X, l2)
        assert np.array_equal(Pc, P)
        assert np.array_equal(impl.pairwise(Pc, l2), D)
        assert np.array_equal(impl.diffab_dense(Pc, l2), diffab)
    Pn = np.array([oracle_np.push_up(parent, length, x, l2) for x in X])
    assert np.array_equal(Pn, P)
    assert np.array_equal(oracle_np.pairwise(Pn, l2), D)
    assert np.array_equal(oracle_np.diffab_dense(Pn, l2), diffab)


def test_toy_known_answers(toy):
    # hand values of the reference's own test: tests/scripts/test_make_all_pw_fununifrac.py:32-33,78-88,173-175
    parent, length, z = toy
    P = oracle_c.push_up(parent, length, z["l1_median.X"], False)
    known = np.array([[0.0, 13 / 14, 10 / 21], [13 / 14, 0.0, 4 / 3], [10 / 21, 4 / 3, 0.0]])
    assert np.allclose(oracle_c.pairwise(P, False), known, atol=1e-3)
    assert np.allclose(oracle_c.diffab_dense(P, False)[0, 2], [-1 / 14, -1 / 21, -5 / 42, 5 / 42, 0, 5 / 42, 0], atol=1e-3)
    P2 = oracle_c.push_up(parent, length, z["l2_median.X"], True)
    known2 = np.array([[0.0, np.sqrt(37) / 14, np.sqrt(22) / 21], [np.sqrt(37) / 14, 0.0, np.sqrt(13) / 6],
                       [np.sqrt(22) / 21, np.sqrt(13) / 6, 0.0]]) ** 2
    assert np.allclose(oracle_c.pairwise(P2, True), known2, atol=1e-3)


@pytest.mark.parametrize("tag,l2,unw", [("l1", False, False), ("l2", True, False), ("l1_unweighted", False, True)])
def test_synth_mid(tag, l2, unw, golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_mid.npz"))
    X = z["X"].copy()
    if unw:
        X[X > 0] = 1
    P = oracle_c.push_up(z["parent"], z["length"], X, l2)
    assert np.array_equal(P, z[f"{tag}.P"])
    D = oracle_c.pairwise(P, l2)
    assert np.array_equal(D, z[f"{tag}.D"])
    assert np.array_equal(oracle_c.pairwise_rows(P, 3, 9, l2), D[3:9])
    # the zero → epsilon rule is exercised by this tree
    assert (z["length"][:-1] == 0).sum() > 0


def test_real_tree_rows(golden_dir, real_trees):
    from fununifrac_b200.factory import make_emd_input, make_tree
    z = np.load(os.path.join(golden_dir, "real_tree_rows.npz"))
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(
        real_trees["kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt"]), "ko00001")
    n, N = len(inp.basis), len(z["indptr"]) - 1
    X = np.zeros((N, n))
    for i in range(N):
        s, e = z["indptr"][i], z["indptr"][i + 1]
        X[i, z["indices"][s:e]] = z["values"][s:e]
    for tag, l2 in [("l1", False), ("l2", True)]:
        P = oracle_c.push_up(inp.parent, inp.edge_length, X, l2)
        assert [sha(p) for p in P] == list(z[f"{tag}.P_sha256"])
        assert np.array_equal(oracle_c.pairwise(P, l2), z[f"{tag}.D"])


def test_solve(golden_dir):
    with open(os.path.join(golden_dir, "solve.json")) as f:
        g = json.load(f)
    with open(os.path.join(golden_dir, "trees.json")) as f:
        trees = json.load(f)
    t4 = trees["small_edge_list_with_lengths_emdu.txt:nobrite"]
    for weighted in (True, False):
        c = g[f"emdu4:weighted={weighted}"]
        Z, diffab = oracle_c.solve(t4["parent"], t4["length"], c["P"], c["Q"])
        assert Z == c["Z"] == (0.25 if weighted else 0.5)  # tests/algorithms/test_emd_unifrac.py:212,248,272
        want = {int(k.split(",")[0]): v for k, v in c["diffab"].items()}
        assert {i: diffab[i] for i in range(4) if diffab[i] != 0} == want
    t7 = trees["small_edge_list_with_lengths.txt:nobrite"]
    for c in g["small7:dirichlet"]:
        Z, diffab = oracle_c.solve(t7["parent"], t7["length"], c["P"], c["Q"])
        assert Z == c["Z"]
        assert np.isclose(Z, c["Z_flow"], atol=1e-12)  # solve_with_flow agrees (LP optimum)
