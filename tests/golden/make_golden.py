#!/usr/bin/env python
"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Run in the authoring container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's hot path (``src/factory/make_tree.py``,
``src/factory/make_emd_input.py``, ``src/algorithms/emd_unifrac.py``,
``src/objects/profile_vector.py``) with two stub modules for the absent
third-party packages (``blist = list``; an empty ``sparse``), feeds it the
reference's own fixtures and seeded synthetic inputs, and writes:

* ``data/``              – copies of the reference's *input* fixtures (toy tree, toy gather CSVs, the
                           two real KO edge lists, gzipped) so the tests do not need ``/root/reference``;
* ``toy_config1.npz``    – BASELINE config 1: X, P_pushed, D, dense diffab for L1 / L2 / unweighted / default key;
* ``trees.json``         – postorder names / parent / length digests for every tree fixture;
* ``synth_mid.npz``      – a 6-level synthetic tree with 24 samples: reference D (L1, L2, unweighted) and P_pushed;
* ``solve.json``         – legacy per-pair ``solve`` / ``solve_with_flow`` outputs;
* ``errors.json``        – exception types the reference raises on malformed inputs.
"""
import contextlib
import glob
import gzip
import hashlib
import io
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

# -- stubs for absent third-party modules -----------------------------------
blist_mod = types.ModuleType("blist")
blist_mod.blist = list
sys.modules["blist"] = blist_mod
sys.modules["sparse"] = types.ModuleType("sparse")
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import pandas as pd  # noqa: E402
import src.factory.make_tree as ref_make_tree  # noqa: E402
import src.factory.make_emd_input as ref_make_emd_input  # noqa: E402
from src.algorithms.emd_unifrac import EarthMoverDistanceUniFracSolver  # noqa: E402
from src.objects.profile_vector import get_L1_diffab, get_L2_diffab  # noqa: E402

from fununifrac_b200 import synth  # noqa: E402  (input generator only; nothing of the product is used as a result)

DATA = os.path.join(HERE, "data")
os.makedirs(DATA, exist_ok=True)


def sha(arr_or_bytes) -> str:
    if isinstance(arr_or_bytes, np.ndarray):
        arr_or_bytes = np.ascontiguousarray(arr_or_bytes).tobytes()
    return hashlib.sha256(arr_or_bytes).hexdigest()


def copy_fixtures():
    small = ["small_edge_list_with_lengths.txt", "small_edge_list_with_lengths_emdu.txt", "small_edge_list.txt",
             "small_sim_10_KOs_gather.csv", "small_sim2_10_KOs_gather.csv", "small_sim3_10_KOs_gather.csv"]
    for f in small:
        shutil.copyfile(os.path.join(REF, "data", f), os.path.join(DATA, f))
    os.makedirs(os.path.join(DATA, "test_input"), exist_ok=True)
    for f in glob.glob(os.path.join(REF, "data", "test_input", "*.txt")):
        shutil.copyfile(f, os.path.join(DATA, "test_input", os.path.basename(f)))
    for f in ["kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt",
              "kegg_tree_scale_10_k_5_deterministic.txt"]:
        with open(os.path.join(REF, "data", f), "rb") as src, \
                gzip.GzipFile(os.path.join(DATA, f + ".gz"), "wb", compresslevel=9, mtime=0) as dst:
            dst.write(src.read())


def emdu_arrays(inp):
    n = len(inp.basis)
    parent = np.full(n, -1, dtype=np.int32)
    length = np.zeros(n, dtype=np.float64)
    for i in range(n - 1):
        parent[i] = inp.Tint[i]
        length[i] = inp.lint[i, inp.Tint[i]]
    names = [inp.idx_to_node[i] for i in range(n)]
    return parent, length, names


def tree_digest(path, brite):
    tree = ref_make_tree.import_graph(path, directed=True)
    inp = ref_make_emd_input.tree_to_EMDU_input(tree, brite)
    parent, length, names = emdu_arrays(inp)
    d = dict(n=len(names), names_sha256=sha("\n".join(names).encode()), parent_sha256=sha(parent),
             length_sha256=sha(length), first=names[:12], last=names[-12:], brite=brite,
             n_lint=len(inp.lint), basis_is_range=(inp.basis == list(range(len(names)))))
    if len(names) <= 64:
        d.update(names=names, parent=parent.tolist(), length=length.tolist())
    return d


def golden_trees():
    out = {}
    out["small_edge_list_with_lengths.txt"] = tree_digest(os.path.join(REF, "data/small_edge_list_with_lengths.txt"), "ko00001")
    out["small_edge_list_with_lengths.txt:nobrite"] = tree_digest(os.path.join(REF, "data/small_edge_list_with_lengths.txt"), None)
    out["small_edge_list_with_lengths_emdu.txt:nobrite"] = tree_digest(os.path.join(REF, "data/small_edge_list_with_lengths_emdu.txt"), None)
    out["test_input/unbalanced_tree.txt:nobrite"] = tree_digest(os.path.join(REF, "data/test_input/unbalanced_tree.txt"), None)
    out["test_input/tree_with_multiple_single_descendants.txt:nobrite"] = tree_digest(
        os.path.join(REF, "data/test_input/tree_with_multiple_single_descendants.txt"), None)
    for f in ["kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt", "kegg_tree_scale_10_k_5_deterministic.txt"]:
        out[f] = tree_digest(os.path.join(REF, "data", f), "ko00001")
    # a sub-brite of a tree WITHOUT a 'root' node exercises set_subtree on an interior node
    out["kegg_tree_scale_10_k_5_deterministic.txt:ko04131"] = tree_digest(
        os.path.join(REF, "data/kegg_tree_scale_10_k_5_deterministic.txt"), "ko04131")
    with open(os.path.join(HERE, "trees.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def run_reference(edge_file, files, brite, key, unweighted, is_L2):
    """The reference driver's compute section (src/compute_all_pairwise_fununifrac.py:51-73), single process."""
    solver = EarthMoverDistanceUniFracSolver()
    tree = ref_make_tree.import_graph(edge_file, directed=True)
    inp = ref_make_emd_input.tree_to_EMDU_input(tree, brite)
    X, PP = [], {}
    for f in files:
        with contextlib.redirect_stdout(io.StringIO()):
            _, P = ref_make_emd_input.get_func_profiles_parallel((f, inp, key, unweighted))
        X.append(P.copy())
        PP[f] = solver.push_up_L2(P, inp) if is_L2 else solver.push_up_L1(P, inp)
    D, _ = solver.pairwise_computation(PP, files, inp, False, is_L2)
    N, n = len(files), len(inp.basis)
    diffab = np.zeros((N, N, n))
    for i in range(N):
        for j in range(i + 1, N):
            d = get_L2_diffab(PP[files[i]], PP[files[j]]) if is_L2 else get_L1_diffab(PP[files[i]], PP[files[j]])
            diffab[i, j] = d
            diffab[j, i] = -d  # emd_unifrac.py:68-72: COO minus its (1,0,2) transpose
    return np.array(X), np.array([PP[f] for f in files]), D, diffab, inp


def golden_toy():
    edge = os.path.join(REF, "data/small_edge_list_with_lengths.txt")
    files = sorted(glob.glob(os.path.join(REF, "data", "*_gather.csv")))
    out = {"files": np.array([os.path.basename(f) for f in files])}
    variants = {"l1_median": ("median_abund", False, False), "l2_median": ("median_abund", False, True),
                "l1_median_unweighted": ("median_abund", True, False), "l2_median_unweighted": ("median_abund", True, True),
                "l1_default": ("f_unique_weighted", False, False)}
    for tag, (key, unw, l2) in variants.items():
        X, PP, D, diffab, inp = run_reference(edge, files, "ko00001", key, unw, l2)
        out[f"{tag}.X"], out[f"{tag}.P"], out[f"{tag}.D"], out[f"{tag}.diffab"] = X, PP, D, diffab
    out["nodes"] = np.array([inp.idx_to_node[i] for i in range(len(inp.basis))])
    np.savez(os.path.join(HERE, "toy_config1.npz"), **out)
    # test_func_profile_convert values (tests/algorithms/test_emd_unifrac.py:89-111)
    tree = ref_make_tree.import_graph(edge, directed=True)
    inp = ref_make_emd_input.tree_to_EMDU_input(tree)
    df = pd.read_csv(os.path.join(REF, "data/small_sim_10_KOs_gather.csv"))
    conv = {}
    for key, norm in [("median_abund", True), ("f_orig_query", True), ("f_orig_query", False)]:
        with contextlib.redirect_stdout(io.StringIO()):
            conv[f"{key}:{norm}"] = ref_make_emd_input.functional_profile_to_vector(df, inp, abundance_key=key, normalize=norm).tolist()
    with open(os.path.join(HERE, "profile_convert.json"), "w") as f:
        json.dump(conv, f, indent=1)


def golden_synth_mid():
    widths = (1, 1, 4, 12, 40, 160, 1200)
    tmp = tempfile.mkdtemp()
    edge = synth.synth_ko_tree(os.path.join(tmp, "synth_tree.txt"), seed=7, widths=widths)
    solver = EarthMoverDistanceUniFracSolver()
    tree = ref_make_tree.import_graph(edge, directed=True)
    inp = ref_make_emd_input.tree_to_EMDU_input(tree, "ko00001")
    parent, length, names = emdu_arrays(inp)
    n, N = len(names), 24
    leaves = synth.leaves_of(parent)
    X = synth.synth_profiles_dense(leaves, n, N, seed=3, nnz_lo=0.05, nnz_hi=0.4)
    # two near-duplicate rows (small D relative to the operand mass) and one exact duplicate
    X[20] = X[0] * (1.0 + 1e-3 * np.cos(np.arange(n)))
    X[20] /= X[20].sum()
    X[21] = X[1]
    out = dict(widths=np.array(widths), tree_seed=7, profile_seed=3, parent=parent, length=length,
               names=np.array(names), X=X)
    files = [f"s{i:03d}" for i in range(N)]
    for tag, l2, unw in [("l1", False, False), ("l2", True, False), ("l1_unweighted", False, True)]:
        PP = {}
        for i, f in enumerate(files):
            P = X[i].copy()
            if unw:
                P[P > 0] = 1
            PP[f] = solver.push_up_L2(P, inp) if l2 else solver.push_up_L1(P, inp)
        D, _ = solver.pairwise_computation(PP, files, inp, False, l2)
        out[f"{tag}.D"] = D
        out[f"{tag}.P"] = np.array([PP[f] for f in files])
    np.savez_compressed(os.path.join(HERE, "synth_mid.npz"), **out)
    shutil.rmtree(tmp)


def golden_real_tree_rows():
    """Reference push-up + distances for 6 seeded samples on the real AAI tree (n = 25,961)."""
    edge = os.path.join(REF, "data/kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt")
    solver = EarthMoverDistanceUniFracSolver()
    tree = ref_make_tree.import_graph(edge, directed=True)
    inp = ref_make_emd_input.tree_to_EMDU_input(tree, "ko00001")
    parent, length, names = emdu_arrays(inp)
    n, N = len(names), 6
    leaves = synth.leaves_of(parent)
    indptr, indices, values = synth.synth_profile_rows(leaves, n, range(N), seed=11)
    X = np.zeros((N, n))
    for i in range(N):
        X[i, indices[indptr[i]:indptr[i + 1]]] = values[indptr[i]:indptr[i + 1]]
    files = [f"r{i}" for i in range(N)]
    out = dict(indptr=indptr, indices=indices, values=values, profile_seed=11)
    for tag, l2 in [("l1", False), ("l2", True)]:
        PP = {f: (solver.push_up_L2(X[i].copy(), inp) if l2 else solver.push_up_L1(X[i].copy(), inp))
              for i, f in enumerate(files)}
        D, _ = solver.pairwise_computation(PP, files, inp, False, l2)
        out[f"{tag}.D"] = D
        out[f"{tag}.P_sha256"] = np.array([sha(PP[f]) for f in files])
        out[f"{tag}.P0_internal"] = PP[files[0]][np.setdiff1d(np.arange(n), leaves)]
    np.savez_compressed(os.path.join(HERE, "real_tree_rows.npz"), **out)


def golden_solve():
    out = {}
    solver = EarthMoverDistanceUniFracSolver()
    tree = ref_make_tree.import_graph(os.path.join(REF, "data/small_edge_list_with_lengths_emdu.txt"))
    inp = ref_make_emd_input.tree_to_EMDU_input(tree)
    samples = {'C': (1, 0), 'B': (1, 1), 'A': (0, 0), 'temp0': (0, 1)}
    P = np.array([samples[inp.idx_to_node[i]][0] for i in range(4)], dtype=float)
    Q = np.array([samples[inp.idx_to_node[i]][1] for i in range(4)], dtype=float)
    for weighted in (True, False):
        p, q = (P / 2, Q / 2) if weighted else (P.copy(), Q.copy())
        Z, diffab = solver.solve(inp, p.copy(), q.copy(), weighted=weighted)
        Z2, F, diffab2 = solver.solve_with_flow(inp, p.copy(), q.copy(), weighted=weighted)
        out[f"emdu4:weighted={weighted}"] = dict(
            P=p.tolist(), Q=q.tolist(), Z=float(Z), diffab={f"{a},{b}": float(v) for (a, b), v in diffab.items()},
            Z_flow=float(Z2), F={f"{a},{b}": float(v) for (a, b), v in F.items()},
            diffab_flow={f"{a},{b}": float(v) for (a, b), v in diffab2.items()})
    tree = ref_make_tree.import_graph(os.path.join(REF, "data/small_edge_list_with_lengths.txt"))
    inp = ref_make_emd_input.tree_to_EMDU_input(tree)
    rng = np.random.default_rng(5)
    cases = []
    for _ in range(6):
        p, q = rng.dirichlet(np.ones(7)), rng.dirichlet(np.ones(7))
        Z, diffab = solver.solve(inp, p.copy(), q.copy(), weighted=True)
        Z2, F, diffab2 = solver.solve_with_flow(inp, p.copy(), q.copy(), weighted=True)
        cases.append(dict(P=p.tolist(), Q=q.tolist(), Z=float(Z), Z_flow=float(Z2),
                          diffab={f"{a},{b}": float(v) for (a, b), v in diffab.items()},
                          F={f"{a},{b}": float(v) for (a, b), v in F.items()},
                          diffab_flow={f"{a},{b}": float(v) for (a, b), v in diffab2.items()}))
    out["small7:dirichlet"] = cases
    with open(os.path.join(HERE, "solve.json"), "w") as f:
        json.dump(out, f, indent=1)


def golden_errors():
    tmp = tempfile.mkdtemp()

    def write(name, text):
        p = os.path.join(tmp, name)
        with open(p, "w") as f:
            f.write(text)
        return p

    cases = {
        "missing_file": lambda: ref_make_tree.import_graph(os.path.join(tmp, "nope.txt")),
        "four_column_header": lambda: ref_make_tree.import_graph(write("h4.txt", "a\tb\tc\td\nx\ty\t1.0\n")),
        "non_float_length": lambda: ref_make_tree.import_graph(write("nf.txt", "p\tc\tlen\nx\ty\tabc\n")),
        "trailing_tab": lambda: ref_make_tree.import_graph(os.path.join(REF, "data/test_input/tree_with_one_single_child.txt")),
        "two_parents": lambda: ref_make_emd_input.tree_to_EMDU_input(
            ref_make_tree.import_graph(write("tp.txt", "p\tc\tlen\nr\ta\t1\nr\tb\t1\na\tc\t1\nb\tc\t1\n"))),
        "two_roots": lambda: ref_make_emd_input.tree_to_EMDU_input(
            ref_make_tree.import_graph(write("tr.txt", "p\tc\tlen\nr\ta\t1\ns\tb\t1\n"))),
        # the AAI tree has a literal 'root' node: a sub-brite then has two in-degree-0 nodes
        "sub_brite_of_tree_with_root_node": lambda: ref_make_emd_input.tree_to_EMDU_input(
            ref_make_tree.import_graph(os.path.join(
                REF, "data/kegg_ko_edge_df_br_ko00001.txt_AAI_lengths_n_50_f_10_r_100.txt")), "ko04131"),
        "unknown_brite_node": lambda: ref_make_emd_input.tree_to_EMDU_input(
            ref_make_tree.import_graph(os.path.join(REF, "data/small_edge_list_with_lengths.txt")), "ko99999"),
        "headerless_loses_first_edge": lambda: sorted(ref_make_emd_input.tree_to_EMDU_input(
            ref_make_tree.import_graph(write("hl.txt", "r\ta\t1.0\nr\tb\t2.0\nb\tc\t3.0\n"))).idx_to_node.values()),
        "two_column_tree_has_no_length": lambda: ref_make_emd_input.tree_to_EMDU_input(
            ref_make_tree.import_graph(os.path.join(REF, "data/small_edge_list.txt"))),
    }
    tree = ref_make_tree.import_graph(os.path.join(REF, "data/small_edge_list_with_lengths.txt"))
    inp = ref_make_emd_input.tree_to_EMDU_input(tree)
    df = pd.read_csv(os.path.join(REF, "data/small_sim_10_KOs_gather.csv"))
    cases["bad_abundance_key_string_column"] = lambda: ref_make_emd_input.functional_profile_to_vector(df, inp, abundance_key="name")
    cases["missing_abundance_column"] = lambda: ref_make_emd_input.functional_profile_to_vector(df, inp, abundance_key="nope")
    cases["all_zero_profile"] = lambda: ref_make_emd_input.functional_profile_to_vector(df, inp, abundance_key="gather_result_rank"
                                                                                         ) if False else ref_make_emd_input.functional_profile_to_vector(df.assign(z=0.0), inp, abundance_key="z")
    cases["nan_name"] = lambda: ref_make_emd_input.functional_profile_to_vector(df.assign(name=np.nan), inp)
    cases["duplicate_ko_last_wins"] = lambda: ref_make_emd_input.functional_profile_to_vector(
        pd.DataFrame({"name": ["ko:d", "ko:e", "x|y|ko:d"], "median_abund": [1.0, 1.0, 3.0]}), inp).tolist()
    cases["unknown_ko_skipped"] = lambda: ref_make_emd_input.functional_profile_to_vector(
        pd.DataFrame({"name": ["ko:d", "ko:ZZZ"], "median_abund": [1.0, 5.0]}), inp).tolist()
    cases["internal_node_accepted"] = lambda: ref_make_emd_input.functional_profile_to_vector(
        pd.DataFrame({"name": ["ko:b", "ko:f"], "median_abund": [1, 3]}), inp).tolist()
    out = {}
    for name, fn in cases.items():
        try:
            with contextlib.redirect_stdout(io.StringIO()) as so:
                res = fn()
            out[name] = dict(raises=None, result=res if isinstance(res, list) else None, stdout=so.getvalue())
        except BaseException as e:  # noqa: BLE001 - we are recording what the reference does
            out[name] = dict(raises=type(e).__name__, bases=[c.__name__ for c in type(e).__mro__[:-2]])
    with open(os.path.join(HERE, "errors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    shutil.rmtree(tmp)


def golden_combined():
    """fununifrac/compute_fununifrac_combined.py: one KO x sample table instead of gather files.  The table is built
    from the three toy gather CSVs (median_abund) plus a KO that is not in the tree and an internal-node row; the
    reference's own make_fununifrac_inputs / push_up_L1 / pairwise_computation produce the expected matrices —
    including its quirk of pushing with L1 weights even when --L2 is given (:57-67)."""
    sys.path.insert(0, os.path.join(REF, "fununifrac"))
    import compute_fununifrac_combined as ref_comb
    edge = os.path.join(REF, "data/small_edge_list_with_lengths.txt")
    files = sorted(glob.glob(os.path.join(REF, "data", "*_gather.csv")))
    cols = {}
    for f in files:
        df = pd.read_csv(f)
        cols[os.path.basename(f).replace("_10_KOs_gather.csv", "")] = {
            str(nm).split(":")[-1]: float(v) for nm, v in zip(df["name"], df["median_abund"])}
    kos = sorted(set().union(*[set(c) for c in cols.values()])) + ["K99999_not_in_tree", "b"]
    table = pd.DataFrame({c: [cols[c].get(k, 0.0) for k in kos] for c in cols}, index=pd.Index(kos, name="name"))
    table.loc["K99999_not_in_tree"] = [1.0, 0.0, 2.0]      # counts in the normalisation, then skipped with a warning
    table.loc["b"] = [0.0, 3.0, 0.0]                        # mass placed on an internal node
    table.to_csv(os.path.join(DATA, "small_combined_table.csv"))
    solver = EarthMoverDistanceUniFracSolver()
    tree = ref_make_tree.import_graph(edge, directed=True)
    inp = ref_make_emd_input.tree_to_EMDU_input(tree, "ko00001")
    sample_df = pd.read_csv(os.path.join(DATA, "small_combined_table.csv"), index_col="name")
    X, PP = [], {}
    for col in sample_df.columns:
        with contextlib.redirect_stdout(io.StringIO()):
            P = ref_comb.make_fununifrac_inputs(sample_df[col], inp)
            assert np.array_equal(P, ref_make_emd_input.extend_vector(sample_df[col], inp))
        X.append(P.copy())
        PP[col] = solver.push_up_L1(P, inp)
    out = {"columns": np.array(list(sample_df.columns)), "X": np.array(X),
           "P": np.array([PP[c] for c in sample_df.columns])}
    out["D_l1"], _ = solver.pairwise_computation(PP, sample_df.columns, inp, False, False)
    out["D_l2_quirk"], _ = solver.pairwise_computation(PP, sample_df.columns, inp, False, True)
    np.savez(os.path.join(HERE, "combined.npz"), **out)


if __name__ == "__main__":
    copy_fixtures()
    golden_combined()
    golden_trees()
    golden_toy()
    golden_synth_mid()
    golden_real_tree_rows()
    golden_solve()
    golden_errors()
    print("golden vectors written to", HERE)
