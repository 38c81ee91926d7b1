"""Host ingest, profile side: gather CSV → float64 vector, bit-exact against the reference's own output
(tests/golden/make_golden.py → toy_config1.npz, profile_convert.json, errors.json)."""
import glob
import json
import os

import numpy as np
import pandas as pd
import pytest

from fununifrac_b200.factory import make_emd_input, make_tree


@pytest.fixture(scope="module")
def toy(data_dir):
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"))
    return make_emd_input.tree_to_EMDU_input(tree)


def test_func_profile_convert(toy, data_dir, golden_dir):
    # reference tests/algorithms/test_emd_unifrac.py:89-111
    df = pd.read_csv(os.path.join(data_dir, "small_sim_10_KOs_gather.csv"))
    P = make_emd_input.functional_profile_to_vector(df, toy, abundance_key='median_abund', normalize=True)
    idx = toy.node_to_idx()
    assert np.isclose(np.sum(P), 1.0, atol=1e-8) and np.all(P >= 0) and np.all(P <= 1)
    assert np.isclose(P[idx['e']], 2 / 6, atol=1e-8)
    assert np.isclose(P[idx['d']], 3 / 6, atol=1e-8)
    assert np.isclose(P[idx['f']], 1 / 6, atol=1e-8)
    with pytest.raises(ValueError):
        make_emd_input.functional_profile_to_vector(df, toy, abundance_key='name', normalize=True)
    P = make_emd_input.functional_profile_to_vector(df, toy, abundance_key='f_orig_query', normalize=False)
    assert np.isclose(np.sum(P), 1.258921, atol=1e-4)
    with open(os.path.join(golden_dir, "profile_convert.json")) as f:
        gold = json.load(f)
    for k, want in gold.items():
        key, norm = k.split(":")
        got = make_emd_input.functional_profile_to_vector(df, toy, abundance_key=key, normalize=(norm == "True"))
        assert got.tolist() == want  # bit-exact float64


@pytest.mark.parametrize("tag,key,unweighted", [("l1_median", "median_abund", False),
                                                ("l1_median_unweighted", "median_abund", True),
                                                ("l1_default", "f_unique_weighted", False)])
def test_toy_profiles_bit_exact(tag, key, unweighted, toy, data_dir, golden_dir):
    z = np.load(os.path.join(golden_dir, "toy_config1.npz"))
    files = sorted(glob.glob(os.path.join(data_dir, "*_gather.csv")))
    assert [os.path.basename(f) for f in files] == list(z["files"])  # sorted-glob basis order: sim2, sim3, sim
    for threads in (1, 2):
        res = make_emd_input.parallel_functional_profile_to_vector(threads, files, toy, key, unweighted)
        assert [f for f, _ in res] == files
        got_files, X = make_emd_input.profiles_to_matrix(res, len(toy.basis))
        assert got_files == files
        assert np.array_equal(X, z[f"{tag}.X"])


def test_profile_edge_cases(toy, data_dir, golden_dir, capsys):
    with open(os.path.join(golden_dir, "errors.json")) as f:
        gold = json.load(f)
    df = pd.read_csv(os.path.join(data_dir, "small_sim_10_KOs_gather.csv"))
    f2v = make_emd_input.functional_profile_to_vector
    with pytest.raises(KeyError):
        f2v(df, toy, abundance_key="nope")
    with pytest.raises(Exception) as ei:
        f2v(df.assign(z=0.0), toy, abundance_key="z")
    assert type(ei.value) is Exception and "empty" in str(ei.value)
    with pytest.raises(Exception) as ei:
        f2v(df.assign(name=np.nan), toy)
    assert type(ei.value) is Exception and "Could not parse the name" in str(ei.value)
    assert f2v(pd.DataFrame({"name": ["ko:d", "ko:e", "x|y|ko:d"], "median_abund": [1.0, 1.0, 3.0]}), toy
               ).tolist() == gold["duplicate_ko_last_wins"]["result"]
    capsys.readouterr()
    assert f2v(pd.DataFrame({"name": ["ko:d", "ko:ZZZ"], "median_abund": [1.0, 5.0]}), toy
               ).tolist() == gold["unknown_ko_skipped"]["result"]
    assert "Warning: ZZZ not found in EMDU index, skipping." in capsys.readouterr().out
    assert f2v(pd.DataFrame({"name": ["ko:b", "ko:f"], "median_abund": [1, 3]}), toy
               ).tolist() == gold["internal_node_accepted"]["result"]
    # empty frame → same "empty profile" Exception as an all-zero one
    with pytest.raises(Exception):
        f2v(df.iloc[0:0], toy)


def test_extend_vector(toy):
    s = pd.Series({"d": 2.0, "f": 6.0, "nope": 1.0})
    P = make_emd_input.extend_vector(s, toy)
    idx = toy.node_to_idx()
    assert P[idx["d"]] == 2.0 / 9.0 and P[idx["f"]] == 6.0 / 9.0 and P.sum() == pytest.approx(8 / 9)


def test_combined_table_to_matrix(data_dir, golden_dir, capsys):
    """Combined KO x sample table (fununifrac/compute_fununifrac_combined.py:13-27, make_emd_input.extend_vector):
    the vectorised table reader and the per-column function against vectors produced by the reference."""
    import pandas as pd
    from fununifrac_b200 import compute_fununifrac_combined as comb
    z = np.load(os.path.join(golden_dir, "combined.npz"))
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"))
    inp = make_emd_input.tree_to_EMDU_input(tree, "ko00001")
    df = pd.read_csv(os.path.join(data_dir, "small_combined_table.csv"), index_col="name")
    assert list(df.columns) == list(z["columns"])
    X = comb.table_to_matrix(df, inp)
    assert "K99999_not_in_tree not found in EMDU index" in capsys.readouterr().out
    assert np.array_equal(X, z["X"])
    for i, col in enumerate(df.columns):
        assert np.array_equal(comb.make_fununifrac_inputs(df[col], inp), z["X"][i])
        assert np.array_equal(make_emd_input.extend_vector(df[col], inp), z["X"][i])
    assert not np.isclose(X.sum(1), 1).all()      # the KO outside the tree took part in the normalisation
