"""Native gather-CSV ingest (csrc/ingest.cu, host code) against pandas + the Python mirror of the reference's
``functional_profile_to_vector`` (itself pinned bit-for-bit to reference-generated goldens in test_host_profiles.py):
rows must be bit-identical, warnings and exceptions the reference's.  Runs on CPU: the library loads without a GPU."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

from fununifrac_b200 import synth
from fununifrac_b200.factory import make_emd_input, make_tree

COLS = synth.GATHER_HEADER.split(",")


@pytest.fixture(scope="module")
def toy(data_dir):
    tree = make_tree.import_graph(os.path.join(data_dir, "small_edge_list_with_lengths.txt"))
    return make_emd_input.tree_to_EMDU_input(tree, "ko00001")


def python_rows(files, inp, key, unweighted=False):
    res = make_emd_input.parallel_functional_profile_to_vector(1, files, inp, key, unweighted)
    return make_emd_input.profiles_to_matrix(res, len(inp.basis))[1]


def write_csv(path, rows, key="median_abund", newline="\n", header=None):
    """rows: (name_field_text, abundance_field_text) already formatted as they should appear in the file."""
    cols = header or COLS
    with open(path, "w", newline="") as f:
        f.write(",".join(cols) + newline)
        for name, ab in rows:
            rec = {c: "0" for c in cols}
            rec.update(name=name, filename="x.sig", md5="ab", query_filename="q.fa", query_name="", moltype="DNA")
            rec[key] = ab
            f.write(",".join(rec[c] for c in cols) + newline)
    return path


def test_toy_files_all_keys(toy, data_dir, capsys):
    files = sorted(glob.glob(os.path.join(data_dir, "*_gather.csv")))
    for key in ("median_abund", "f_unique_weighted", "average_abund", "f_orig_query", "std_abund", "intersect_bp"):
        for unw in (False, True):
            want = python_rows(files, toy, key, unw)
            got_files, got = make_emd_input.native_profiles_to_matrix(files, toy, key, unw, 2)
            assert got_files == files and np.array_equal(got, want), (key, unw)
    capsys.readouterr()


def test_float_formats_bit_identical_to_pandas(toy, tmp_path):
    """pandas' default C float parser is not strtod: digits beyond 17 are dropped, one multiply/divide by a power of ten."""
    rng = np.random.default_rng(0)
    texts = ["3", "3.0", "+2.5", "0.683218948", "1e-3", "1E+2", "2.5e-05", "123456789012345678", "0.1234567890123456789012",
             "1.7976931348623157e308", "4.9e-324", "2.2250738585072014e-308", "1e-400", "0.000000000000000000001", " 7 ", "7.",
             ".5", "00012.50", "9007199254740993", "1e22", "1e23", "8.41e21", "0.30000000000000004", "5e-1"]
    texts += [repr(float(x)) for x in rng.random(300)]
    texts += [f"{x:.{int(d)}g}" for x, d in zip(10 ** rng.uniform(-12, 12, 300), rng.integers(1, 20, 300))]
    texts += [f"{x:.{int(d)}e}" for x, d in zip(10 ** rng.uniform(-300, 300, 200), rng.integers(0, 19, 200))]
    nodes = ["d", "e", "f", "g"]
    for chunk in range(0, len(texts), 4):
        part = texts[chunk:chunk + 4]
        rows = [(f"ko:{nodes[i]}", t) for i, t in enumerate(part)]
        path = write_csv(str(tmp_path / f"f{chunk}_gather.csv"), rows)
        df = pd.read_csv(path)
        assert df["median_abund"].dtype.kind in "fi", part           # pandas sees numbers
        want = python_rows([path], toy, "median_abund")
        got = make_emd_input.native_profiles_to_matrix([path], toy, "median_abund", False, 1)[1]
        assert np.array_equal(got, want), part


def test_names_quotes_duplicates_unknown_ids(toy, tmp_path, capsys):
    rows = [('"abc|x, with a comma|ko:d"', "3"), ("ko:e", "2"), ('"say ""hi""|ko:f"', "1.5"), ("lmn|opq|ko:g", "4"),
            ("ko:K99999", "8"), ("prefix|ko:e", "5"), ("b", "0.25"), ("ko:d", "7"), ("x:y:ko00001", "0.5")]
    for nl in ("\n", "\r\n"):
        path = write_csv(str(tmp_path / f"q{len(nl)}_gather.csv"), rows, newline=nl)
        with open(path, "a") as f:
            f.write(nl + nl)                                         # trailing blank lines are skipped
        want = python_rows([path], toy, "median_abund")
        out_py = capsys.readouterr().out
        got = make_emd_input.native_profiles_to_matrix([path], toy, "median_abund", False, 1)[1]
        out_nat = capsys.readouterr().out
        assert np.array_equal(got, want) and out_py == out_nat and "K99999 not found in EMDU index" in out_nat
        idx = toy.node_to_idx()
        assert got[0, idx["d"]] * want.sum() > 0 and got[0, idx["e"]] == want[0, idx["e"]]


def test_errors_like_the_reference(toy, tmp_path, data_dir, capsys):
    good = os.path.join(data_dir, "small_sim_10_KOs_gather.csv")
    bad_num = write_csv(str(tmp_path / "bad_gather.csv"), [("ko:d", "3"), ("ko:e", "abc")])
    with pytest.raises(ValueError, match="gives a value that is not a number: abc"):
        make_emd_input.native_profiles_to_matrix([good, bad_num], toy, "median_abund", False, 2)
    with pytest.raises(ValueError):
        python_rows([bad_num], toy, "median_abund")
    empty = write_csv(str(tmp_path / "zero_gather.csv"), [("ko:d", "0"), ("ko:K99999", "5")])
    with pytest.raises(Exception, match="Functional profile is empty"):
        make_emd_input.native_profiles_to_matrix([empty], toy, "median_abund", False, 1)
    with pytest.raises(Exception, match="Functional profile is empty"):
        python_rows([empty], toy, "median_abund")
    with pytest.raises(KeyError):
        make_emd_input.native_profiles_to_matrix([good], toy, "no_such_key", False, 1)
    with pytest.raises(KeyError):
        python_rows([good], toy, "no_such_key")
    with pytest.raises(FileNotFoundError):       # re-read through pandas: the reference's own exception
        make_emd_input.native_profiles_to_matrix([str(tmp_path / "missing.csv")], toy, "median_abund", False, 1)
    # NA abundance -> NaN propagates exactly as in the reference (nan sum is not == 0, so no "empty" error)
    na = write_csv(str(tmp_path / "na_gather.csv"), [("ko:d", "3"), ("ko:e", "NA"), ("ko:f", "")])
    want = python_rows([na], toy, "median_abund")
    got = make_emd_input.native_profiles_to_matrix([na], toy, "median_abund", False, 1)[1]
    assert np.isnan(want).all() and np.isnan(got).all()
    capsys.readouterr()


def _outcome(fn):
    try:
        return "ok", fn()
    except Exception as e:  # noqa: BLE001
        return type(e).__name__, str(e)


def test_odd_inputs_behave_like_the_pandas_route(toy, tmp_path, capsys):
    """Inputs the native tokeniser is stricter on than pandas (all-bool abundances, '1_0', overflowing exponents, an
    extra field per row = pandas' implicit index column, numeric names): the file is re-read through the reference's
    route, so the observable outcome — the row, or the exception type and message — is the same."""
    cases = {
        "bool": [("ko:d", "True"), ("ko:e", "False")],
        "underscore": [("ko:d", "1_0"), ("ko:e", "abc")],
        "overflow": [("ko:d", "1e400"), ("ko:e", "2")],
        "numeric_names": [("17", "3"), ("42", "2")],
    }
    for tag, rows in cases.items():
        path = write_csv(str(tmp_path / f"{tag}_gather.csv"), rows)
        kind_py, val_py = _outcome(lambda: python_rows([path], toy, "median_abund"))
        kind_nat, val_nat = _outcome(lambda: make_emd_input.native_profiles_to_matrix([path], toy, "median_abund", False, 1)[1])
        assert kind_py == kind_nat, (tag, kind_py, val_py, kind_nat, val_nat)
        if kind_py == "ok":
            assert np.array_equal(val_py, val_nat, equal_nan=True), tag
        else:
            assert val_py == val_nat, tag
    # one extra field in every data row: pandas takes the first column as the index
    path = write_csv(str(tmp_path / "extra_gather.csv"), [("ko:d", "3"), ("ko:e", "2")])
    lines = open(path).read().splitlines()
    with open(path, "w") as f:
        f.write("\n".join([lines[0]] + [ln + ",9" for ln in lines[1:]]) + "\n")
    kind_py, val_py = _outcome(lambda: python_rows([path], toy, "median_abund"))
    kind_nat, val_nat = _outcome(lambda: make_emd_input.native_profiles_to_matrix([path], toy, "median_abund", False, 1)[1])
    assert kind_py == kind_nat and (kind_py != "ok" or np.array_equal(val_py, val_nat, equal_nan=True))
    capsys.readouterr()


def test_many_files_threads_and_order(tmp_path, capsys):
    """A few hundred synthetic gather files on a mid-size synthetic tree: row order = file order, any thread count."""
    path = synth.synth_ko_tree(str(tmp_path / "tree.txt"), seed=3, widths=(1, 1, 4, 12, 40, 160, 900))
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
    n = len(inp.basis)
    leaves = synth.leaves_of(inp.parent)
    names = [inp.idx_to_node[k] for k in range(n)]
    files = []
    rng = np.random.default_rng(5)
    for i in range(120):
        indptr, indices, values = synth.synth_profile_rows(leaves, n, [i], seed=9)
        ab = rng.geometric(0.2, len(indices)) + rng.random(len(indices)).round(int(rng.integers(0, 12)))
        files.append(synth.write_gather_csv(str(tmp_path / f"s{i:04d}_gather.csv"), [names[j] for j in indices], ab))
    want = python_rows(files, inp, "median_abund")
    for threads in (1, 3, 0):
        got = make_emd_input.native_profiles_to_matrix(files, inp, "median_abund", False, threads)[1]
        assert np.array_equal(got, want)
    out = np.full((130, n + 5), -1.0)
    make_emd_input.native_profiles_to_matrix(files, inp, "f_unique_weighted", True, 4, out=out)
    assert np.array_equal(out[:120, :n], python_rows(files, inp, "f_unique_weighted", True)) and np.all(out[120:] == -1)
    capsys.readouterr()


def test_fast_row_scanner_matches_pandas_on_random_layouts(toy, tmp_path, capsys):
    """The data rows go through a 32-bytes-at-a-time scanner (AVX2) unless a record holds a quote character.  Random
    layouts exercise what that scanner must get right: records of every length around the 32-byte blocks, '\\n', '\\r\\n' and
    lone '\\r' terminators, a missing final newline, blank lines, a trailing comma at the end of the input, quoted fields
    (with commas, doubled quotes) in the wanted and in other columns, files shorter than one block — against pandas + the
    Python mirror, bit for bit; the same files once more with the scanner switched off (subprocess) must give the same rows."""
    import subprocess
    import sys
    rng = np.random.default_rng(5)
    nodes = ["d", "e", "f", "g", "b", "c"]
    files = []
    for t in range(60):
        width = int(rng.integers(2, 40))                    # columns: name, median_abund and filler columns
        pos = rng.choice(width, 2, replace=False)
        cols = [f"c{i}" for i in range(width)]
        cols[pos[0]], cols[pos[1]] = "name", "median_abund"
        nl = ["\n", "\r\n", "\r"][t % 3] if t % 7 else "\n"
        lines = [",".join(cols)]
        for r in range(int(rng.integers(1, 12))):
            rec = []
            for c in cols:
                if c == "name":
                    nm = f"ko:{nodes[int(rng.integers(len(nodes)))]}"
                    style = int(rng.integers(5))
                    rec.append([nm, f"x|{nm}", f'"{nm}"', f'"pre, with comma|{nm}"', f'"say ""hi""|{nm}"'][style])
                elif c == "median_abund":
                    rec.append(["3", "2.5", "1e-3", "", "NA", "0.125", "17.0"][int(rng.integers(7))])
                else:
                    k = int(rng.integers(0, 14))
                    filler = "".join(rng.choice(list("abcxyz0123456789. "), k))
                    rec.append(f'"{filler},q"' if rng.random() < 0.05 else filler)
            lines.append(",".join(rec))
            if rng.random() < 0.15:
                lines.append("")                             # blank line
        text = nl.join(lines)
        if t % 4:
            text += nl                                       # (else: no final newline)
        if t % 11 == 0 and nl == "\n":
            text += ",".join([""] * width)                   # last record: only delimiters, ends on a comma at end of input
        path = str(tmp_path / f"r{t:02d}_gather.csv")
        with open(path, "w", newline="") as f:
            f.write(text)
        files.append(path)
    good, want = [], []
    for path in files:
        try:
            row = python_rows([path], toy, "median_abund")
        except Exception:
            # the pandas route rejects the file (e.g. an all-NA profile): the native route must refuse it too
            with pytest.raises(Exception):
                make_emd_input.native_profiles_to_matrix([path], toy, "median_abund", False, 1)
            continue
        good.append(path)
        want.append(row[0])
    capsys.readouterr()
    assert len(good) >= 30
    got = make_emd_input.native_profiles_to_matrix(good, toy, "median_abund", False, 2)[1]
    assert np.array_equal(got, np.stack(want), equal_nan=True)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from fununifrac_b200.factory import make_emd_input, make_tree; "
            "inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(%r), 'ko00001'); "
            "X = make_emd_input.native_profiles_to_matrix(%r, inp, 'median_abund', False, 1)[1]; np.save(%r, X)")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    edge = os.path.join(repo, "tests", "golden", "data", "small_edge_list_with_lengths.txt")
    out = str(tmp_path / "slow.npy")
    subprocess.run([sys.executable, "-c", code % (repo, edge, good, out)], check=True, capture_output=True,
                   env=dict(os.environ, FUF_INGEST_NO_AVX2="1"))
    assert np.array_equal(np.load(out), got, equal_nan=True)
