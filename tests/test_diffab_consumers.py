"""Consumers of --diffab output: ``DiffabArrayIndexer`` / ``convert_diffab_array_to_df`` restated from
the reference's own test ``tests/algorithms/test_emd_unifrac.py:142-184``, on the diffab tensors the unmodified reference
produced for the toy files (tests/golden/toy_config1.npz), held dense and as the COO that ``*.diffab.npz`` loads into.
The GPU-resident store is covered in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from fununifrac_b200 import sparse_npz
from fununifrac_b200.factory import make_emd_input, make_tree
from fununifrac_b200.utility import differential_abundance as da


@pytest.fixture(scope="module")
def toy(data_dir, golden_dir):
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"), directed=True)
    inp = make_emd_input.tree_to_EMDU_input(tree, "ko00001")
    return inp, np.load(os.path.join(golden_dir, "toy_config1.npz"), allow_pickle=True)


def check_indexer(indexer, diffabs, files, atol=1e-8):
    """The assertions of the reference's test_diffab_indexer, for any kind of store."""
    n = diffabs.shape[2]
    assert np.allclose(indexer.get_diffab(files[0], files[0]), np.zeros(n), atol=atol)
    assert np.allclose(indexer.get_diffab(files[1], files[1]), np.zeros(n), atol=atol)
    assert np.allclose(indexer.get_diffab(files[1], files[0]), diffabs[1, 0], atol=atol)
    assert np.allclose(indexer.get_diffab(files[0], files[1]), diffabs[0, 1], atol=atol)
    assert indexer.get_diffab(files[0], files[1]).shape == (1, 1, n)
    assert np.allclose(indexer.get_diffab(files[0], files), diffabs[[0], :, :], atol=atol)
    assert np.allclose(indexer.get_diffab(files, files[1]), diffabs[:, [1], :], atol=atol)
    assert np.allclose(indexer.get_diffab(files, files), diffabs, atol=atol)
    order = ['d', 'e', 'b', 'f', 'g', 'c', 'ko00001']
    assert np.allclose(indexer.get_diffab_for_node(files[0], files[1], order), diffabs[0, 1, :], atol=atol)
    assert np.allclose(indexer.get_diffab_for_node(files[0], files[1], 'ko00001'), diffabs[0, 1, -1], atol=atol)
    assert np.allclose(indexer.get_diffab_for_node(files[0], files, 'd'), diffabs[0][:, [0]], atol=atol)
    got = indexer.get_diffab_for_node(files[0], files, ['ko00001', 'g'])
    assert got.shape == (1, len(files), 2) and np.allclose(got, diffabs[0][:, [-1, 4]], atol=atol)
    # repeated and reordered files
    got = indexer.get_diffab([files[2], files[0], files[2]], [files[1], files[1]])
    assert np.allclose(got, diffabs[np.ix_([2, 0, 2], [1, 1], np.arange(n))], atol=atol)
    with pytest.raises(KeyError):
        indexer.get_diffab("no_such_file", files[0])
    with pytest.raises(KeyError):
        indexer.get_diffab_for_node(files[0], files[1], "K99999")


@pytest.mark.parametrize("tag", ["l1_median", "l2_median", "l1_median_unweighted"])
def test_indexer_dense_and_coo(toy, tag):
    inp, z = toy
    files = [str(f) for f in z["files"]]
    diffabs = z[f"{tag}.diffab"]
    assert list(inp.basis) == list(range(7)) and [inp.idx_to_node[i] for i in range(7)] == [str(x) for x in z["nodes"]]
    check_indexer(da.DiffabArrayIndexer(diffabs, inp.basis, files, inp.idx_to_node), diffabs, files)
    i, j, k = np.nonzero(diffabs)
    coo = sparse_npz.COO(np.stack([i, j, k]), diffabs[i, j, k], diffabs.shape)
    check_indexer(da.DiffabArrayIndexer(coo, inp.basis, files, inp.idx_to_node), diffabs, files, atol=0)


def test_coo_take_matches_ix(tmp_path):
    rng = np.random.default_rng(3)
    dense = rng.normal(size=(6, 6, 11)) * (rng.random((6, 6, 11)) < 0.3)
    i, j, k = np.nonzero(dense)
    coo = sparse_npz.COO(np.stack([i, j, k]), dense[i, j, k], dense.shape)
    sparse_npz.save_npz(str(tmp_path / "t.npz"), coo)
    coo = sparse_npz.load_npz(str(tmp_path / "t.npz"))
    for _ in range(20):
        idx = [rng.integers(-s, s, size=rng.integers(0, 5)) for s in dense.shape]
        assert np.array_equal(coo.take(*idx), dense[np.ix_(*[np.where(a < 0, a + s, a) for a, s in zip(idx, dense.shape)])])
    with pytest.raises(IndexError):
        coo.take([6], [0], [0])
    with pytest.raises(IndexError):
        coo.take([0], [0])


def test_convert_diffab_array_to_df(toy):
    inp, z = toy
    files = ["/some/dir/" + str(f) for f in z["files"]]
    diffabs = z["l1_median.diffab"]
    nodes = [str(x) for x in z["nodes"]]
    df = da.convert_diffab_array_to_df(diffabs, nodes, files)
    base = [os.path.basename(f) for f in files]
    assert list(df.columns) == base and list(df.index) == base
    for i, f in enumerate(base):
        for j, g in enumerate(base):
            cell = df[f][g]
            assert list(cell) == nodes and all(cell[nd] == diffabs[i, j, k] for k, nd in enumerate(nodes))
    with pytest.raises(ValueError, match="Number of samples does not match"):
        da.convert_diffab_array_to_df(diffabs, nodes, files[:2])
    with pytest.raises(ValueError, match="not square"):
        da.convert_diffab_array_to_df(diffabs[:, :2], nodes, files)
    with pytest.raises(ValueError, match="does not match nodes in order"):
        da.convert_diffab_array_to_df(diffabs, nodes[:-1], files)
