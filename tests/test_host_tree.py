"""Host ingest, tree side: bit-exact postorder / parent / lengths against vectors produced by the
reference itself (tests/golden/make_golden.py → trees.json, errors.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

from fununifrac_b200.factory import make_emd_input, make_tree
from fununifrac_b200.objects.func_tree import FuncTree, GraphError


def sha(x):
    if isinstance(x, np.ndarray):
        x = np.ascontiguousarray(x).tobytes()
    return hashlib.sha256(x).hexdigest()


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "trees.json")) as f:
        return json.load(f)


def _check(inp, g):
    n = len(inp.basis)
    names = [inp.idx_to_node[i] for i in range(n)]
    assert n == g["n"]
    assert inp.basis == list(range(n))
    assert names[:12] == g["first"] and names[-12:] == g["last"]
    assert sha("\n".join(names).encode()) == g["names_sha256"]
    parent, length = inp.dense_arrays()
    assert sha(parent) == g["parent_sha256"]
    assert sha(length) == g["length_sha256"]          # float64 bit patterns
    assert np.array_equal(parent, inp.parent) and np.array_equal(length, inp.edge_length)
    assert len(inp.lint) == g["n_lint"]
    # postorder invariants the kernels rely on
    assert np.all(parent[:-1] > np.arange(n - 1)) and parent[-1] == -1
    if "names" in g:
        assert names == g["names"] and parent.tolist() == g["parent"] and length.tolist() == g["length"]


@pytest.mark.parametrize("key", [
    "small_edge_list_with_lengths.txt", "small_edge_list_with_lengths.txt:nobrite",
    "small_edge_list_with_lengths_emdu.txt:nobrite", "test_input/unbalanced_tree.txt:nobrite",
    "test_input/tree_with_multiple_single_descendants.txt:nobrite"])
def test_small_trees(key, golden, data_dir):
    fname = key.split(":")[0]
    brite = golden[key]["brite"]
    tree = make_tree.import_graph(os.path.join(data_dir, fname), directed=True)
    _check(make_emd_input.tree_to_EMDU_input(tree, brite), golden[key])


@pytest.mark.parametrize("key", [
    "kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt",
    "kegg_tree_scale_10_k_5_deterministic.txt",
    "kegg_tree_scale_10_k_5_deterministic.txt:ko04131"])
def test_real_trees(key, golden, real_trees):
    fname = key.split(":")[0]
    tree = make_tree.import_graph(real_trees[fname], directed=True)
    _check(make_emd_input.tree_to_EMDU_input(tree, golden[key]["brite"]), golden[key])


def test_diffab_indexer_postorder(data_dir):
    # reference tests/algorithms/test_emd_unifrac.py:178
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"))
    inp = make_emd_input.tree_to_EMDU_input(tree)
    assert [inp.idx_to_node[i] for i in inp.basis] == ['d', 'e', 'b', 'f', 'g', 'c', 'ko00001']
    assert inp.Tint == {0: 2, 1: 2, 2: 6, 3: 5, 4: 5, 5: 6}
    assert inp.lint[(0, 2)] == inp.lint[(2, 0)] == 0.99993882393976


def test_error_behaviour(golden_dir, data_dir, real_trees, tmp_path):
    with open(os.path.join(golden_dir, "errors.json")) as f:
        gold = json.load(f)

    def write(name, text):
        p = tmp_path / name
        p.write_text(text)
        return str(p)

    cases = {
        "missing_file": lambda: make_tree.import_graph(str(tmp_path / "nope.txt")),
        "four_column_header": lambda: make_tree.import_graph(write("h4.txt", "a\tb\tc\td\nx\ty\t1.0\n")),
        "non_float_length": lambda: make_tree.import_graph(write("nf.txt", "p\tc\tlen\nx\ty\tabc\n")),
        "trailing_tab": lambda: make_tree.import_graph(os.path.join(data_dir, "test_input/tree_with_one_single_child.txt")),
        "two_parents": lambda: make_emd_input.tree_to_EMDU_input(
            make_tree.import_graph(write("tp.txt", "p\tc\tlen\nr\ta\t1\nr\tb\t1\na\tc\t1\nb\tc\t1\n"))),
        "two_roots": lambda: make_emd_input.tree_to_EMDU_input(
            make_tree.import_graph(write("tr.txt", "p\tc\tlen\nr\ta\t1\ns\tb\t1\n"))),
        "sub_brite_of_tree_with_root_node": lambda: make_emd_input.tree_to_EMDU_input(
            make_tree.import_graph(real_trees["kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt"]), "ko04131"),
        "two_column_tree_has_no_length": lambda: make_emd_input.tree_to_EMDU_input(
            make_tree.import_graph(os.path.join(data_dir, "small_edge_list.txt"))),
    }
    for name, fn in cases.items():
        want = gold[name]["raises"]
        with pytest.raises(Exception) as ei:
            fn()
        # same class for builtin exceptions; the reference's generic ones are plain Exception
        assert type(ei.value).__name__ == want, name
    # networkx raises NetworkXError for an unknown brite; ours is GraphError (both direct Exception subclasses)
    assert gold["unknown_brite_node"]["raises"] == "NetworkXError"
    with pytest.raises(GraphError):
        make_emd_input.tree_to_EMDU_input(
            make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt")), "ko99999")
    # a header-less file silently loses its first edge
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(write("hl.txt", "r\ta\t1.0\nr\tb\t2.0\nb\tc\t3.0\n")))
    assert sorted(inp.idx_to_node.values()) == gold["headerless_loses_first_edge"]["result"]


def test_functree_helpers(data_dir):
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"))
    G = tree.main_tree
    assert FuncTree.get_root_of_tree(G) == "ko00001"
    assert FuncTree.get_leaf_nodes(G) == ['d', 'e', 'f', 'g']
    assert FuncTree.get_descendants(G, 'b') == {'b', 'd', 'e'}
    assert FuncTree.get_descendants(G, 'ko00001', leaf_only=True) == {'d', 'e', 'f', 'g'}
    assert FuncTree.get_descendant(G, 'b', 'd') == 'd' and FuncTree.get_descendant(G, 'd', 'b') == 'd'
    assert FuncTree.infer_edge_len_property(G) == 'edge_length'
    with pytest.raises(Exception):
        tree.current_subtree
    tree.set_subtree('c')
    assert tree.basis == ['c', 'f', 'g'] and tree.basis_index == {'c': 0, 'f': 1, 'g': 2}
