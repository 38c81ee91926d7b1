"""BASELINE.json configs 2 and 5 as collected GPU tests (run with ``-m gpu`` on the B200 box), plus the ``--diffab``
CLI on near-duplicate samples.  Everything goes through the drop-in CLI / the C ABI and is checked against the CPU
oracle (``oracle/oracle.c``) on the same inputs.

    python tests/test_gpu_configs.py [N=1000]       # config 2 as a script: prints one JSON line with the timings

Reference tests these mirror: ``tests/scripts/test_make_all_pw_fununifrac.py:10-88`` (the CLI run on gather CSVs, the
saved matrix and the diffab values), ``src/algorithms/emd_unifrac.py:49,56-67`` (pair order and difference rows).
"""
import glob
import itertools
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from fununifrac_b200 import synth  # noqa: E402
from fununifrac_b200.factory import make_emd_input, make_tree  # noqa: E402

pytestmark = pytest.mark.gpu
REL = 1e-6


def _write(args):
    out, leaves, n, names, i = args
    indptr, indices, values = synth.synth_profile_rows(leaves, n, [i], seed=1)
    ab = 1.0 + np.random.default_rng([7, i]).geometric(0.15, len(indices))
    synth.write_gather_csv(os.path.join(out, f"s{i:05d}_gather.csv"), [names[j] for j in indices], ab)
    return len(indices)


def run_config2(N, tmp):
    """Config 2: synthetic full KO tree (30,002 nodes), N gather CSVs on disk (29 sourmash columns, 2,000-8,000 KO rows
    each), weighted L1, median_abund, CLI on one GPU.  Times the whole user-visible run (argument parsing -> tree ->
    native ingest -> push-up + all pairs -> np.save) and checks the saved matrix against the oracle on the same files."""
    from fununifrac_b200 import compute_fununifrac
    from fununifrac_b200.utility import constant
    from oracle import oracle_c
    tree = synth.synth_ko_tree(os.path.join(tmp, "ko_tree.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(tree), "ko00001")
    n = len(inp.basis)
    leaves = synth.leaves_of(inp.parent)
    names = [inp.idx_to_node[k] for k in range(n)]
    prof = os.path.join(tmp, "profiles")
    os.makedirs(prof)
    t = time.perf_counter()
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        rows = sum(ex.map(_write, [(prof, leaves, n, names, i) for i in range(N)], chunksize=8))
    t_gen = time.perf_counter() - t
    files = sorted(glob.glob(os.path.join(prof, "*_gather.csv")))
    nbytes = sum(os.path.getsize(f) for f in files)
    argv = ["-e", tree, "-fd", prof, "-o", os.path.join(tmp, "out"), "-i", "cfg2", "-b", "ko00001", "-a", "median_abund",
            "-t", str(os.cpu_count() or 1)]
    compute_fununifrac.main(argv + ["--force"])                     # first call: CUDA context, library load
    t = time.perf_counter()
    compute_fununifrac.main(argv + ["--force"])
    t_cli = time.perf_counter() - t
    D = np.load(os.path.join(tmp, "out", constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("cfg2")))
    with open(os.path.join(tmp, "out", constant.FUNUNIFRAC_OUT__MAIN_BASIS__FILE_NAME.format("cfg2"))) as f:
        assert [l.strip() for l in f] == files                      # basis = sorted glob order, one path per line
    # oracle on the same files
    t = time.perf_counter()
    _, X = make_emd_input.native_profiles_to_matrix(files, inp, "median_abund", False, os.cpu_count() or 1)
    t_ingest = time.perf_counter() - t
    oracle_c.set_threads(os.cpu_count() or 1)
    t = time.perf_counter()
    Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, False), False)
    t_cpu = time.perf_counter() - t
    assert D.shape == (N, N) and D.dtype == np.float64 and np.array_equal(D, D.T) and not np.diag(D).any()
    iu = np.triu_indices(N, 1)
    err = float((np.abs(D[iu] - Dref[iu]) / Dref[iu]).max())
    return {
        "config": f"BASELINE config 2: {N} gather CSVs ({rows} KO rows, {nbytes / 1e6:.0f} MB), synthetic full KO tree "
                  f"(n={n}), weighted L1, median_abund, CLI on one GPU",
        "cli_seconds": t_cli, "of_which_native_ingest_seconds": t_ingest, "pairs": N * (N - 1) // 2,
        "pairs_per_second_whole_cli": N * (N - 1) // 2 / t_cli, "max_rel_err_vs_oracle": err,
        "cpu_oracle_port_seconds_compute_only": t_cpu, "threads": os.cpu_count(), "csv_generation_seconds": t_gen}


def test_config2_cli_1000_gather_files(tmp_path):
    """BASELINE config 2 at its full size through the CLI; every one of the 499,500 distances within 1e-6 of the oracle."""
    res = run_config2(1000, str(tmp_path))
    assert res["max_rel_err_vs_oracle"] < REL, res
    print(json.dumps(res))


@pytest.fixture(scope="module")
def full_tree(tmp_path_factory):
    from fununifrac_b200 import engine
    path = synth.synth_ko_tree(str(tmp_path_factory.mktemp("synth") / "ko_tree.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
    return path, inp, engine.TreeContext(inp.parent, inp.edge_length), synth.leaves_of(inp.parent)


def test_config5_diffab_stream_500_samples(full_tree):
    """BASELINE config 5: 500 samples on the full tree, every pair's signed difference vector streamed to pinned host
    memory.  The whole stream (124,750 rows x 30,002 entries) is consumed; checked are (a) the pair order — row p is pair
    ``itertools.combinations(range(N), 2)[p]`` — and the values of 600 sampled rows (block edges, row changes of i, random)
    against float64 differences of the oracle's pushed profiles; (b) a checksum of checksums: the sum of all streamed
    rows, per node, equals sum_i (N - 1 - 2 i) P_i[k], evaluated in float64 from the oracle's profiles; (c) the float64
    validation stream is bit-identical to the oracle's P_i - P_j on the sampled rows."""
    import torch

    from fununifrac_b200 import engine
    from oracle import oracle_c
    path, inp, ctx, leaves = full_tree
    N, n = 500, ctx.n
    X = synth.synth_profiles_dense(leaves, n, N, seed=5)
    X[77] = X[3] * (1 + 1e-6 * np.random.default_rng(0).random(n))       # near-duplicates: tiny but non-zero differences
    X[77] /= X[77].sum()
    X[499] = X[4]                                                        # identical samples: an all-zero row
    P = oracle_c.push_up(inp.parent, inp.edge_length, X, False)
    job = ctx.push_up_job(torch.from_numpy(X).cuda(), engine.MODE_L1)
    total = N * (N - 1) // 2
    pairs = list(itertools.combinations(range(N), 2))
    ppb = 1024
    rng = np.random.default_rng(1)
    wanted = set(rng.integers(0, total, 300).tolist())
    wanted |= {0, total - 1, pairs.index((3, 77)), pairs.index((4, 499))}
    wanted |= {p for b in range(0, total, ppb * 16) for p in (b, min(total, b + ppb) - 1)}          # block edges
    wanted |= {engine.combinations_offset(i, N) + d for i in range(1, N - 1, 7) for d in (-1, 0)}   # i changes here
    colsum = np.zeros(n)
    seen, checked = 0, 0
    scale = np.abs(P).max()
    for b, e, rows in engine.iter_diffab_blocks(ctx, job.P64T, N, engine.MODE_L1, pairs_per_block=ppb,
                                                out_dtype=torch.float32):
        assert b == seen and rows.dtype == np.float32 and rows.shape == (e - b, n)
        colsum += rows.sum(0, dtype=np.float64)
        for p in sorted(w for w in wanted if b <= w < e):
            i, j = pairs[p]
            want = P[i] - P[j]
            got = rows[p - b].astype(np.float64)
            # float64 subtraction rounded once to fp32: 6e-8 relative per entry (+ the push-up's own 1e-14)
            assert np.all(np.abs(got - want) <= 6.1e-8 * np.abs(want) + 1e-13 * scale), (p, i, j)
            assert np.array_equal(got != 0, want != 0) or np.abs(want[(got != 0) != (want != 0)]).max() < 1e-13 * scale
            checked += 1
        seen = e
    assert seen == total and checked == len(wanted)
    # (b) every streamed row enters this identity
    coef = (N - 1 - 2 * np.arange(N)).astype(np.float64)
    want_sum = coef @ P
    assert np.all(np.abs(colsum - want_sum) <= 1e-6 * (np.abs(coef)[:, None] * np.abs(P)).sum(0) + 1e-12)
    p_same = pairs.index((4, 499))
    # (c) validation stream: reference operation order in the push-up, float64 rows, bit-identical
    P64T = ctx.push_up_f64(torch.from_numpy(X).cuda(), engine.MODE_L1)
    for p0 in (0, p_same - 5, pairs.index((3, 77)) - 17, total - 40):
        b, e, rows = next(engine.iter_diffab_blocks(ctx, P64T, N, engine.MODE_L1, 40, pair_begin=p0, pair_end=p0 + 40))
        want = np.stack([P[i] - P[j] for i, j in pairs[p0:p0 + 40]])
        assert np.array_equal(rows, want)
    b, e, rows = next(engine.iter_diffab_blocks(ctx, P64T, N, engine.MODE_L1, 1, pair_begin=p_same, pair_end=p_same + 1))
    assert not rows.any()
    # --L2 form of the same stream: squared differences
    b, e, rows = next(engine.iter_diffab_blocks(ctx, P64T, N, engine.MODE_L2, 33, pair_begin=1000, pair_end=1033))
    want = np.stack([(P[i] - P[j]) ** 2 for i, j in pairs[1000:1033]])
    assert np.array_equal(rows, want)


@pytest.mark.parametrize("l2", [False, True])
def test_cli_diffab_near_duplicates_values_and_nnz(l2, full_tree, tmp_path):
    """``--diffab`` through the CLI on the full tree with near-duplicate and identical samples: the saved COO has the
    non-zero pattern of the float64 ``np.nonzero(P - Q)`` (reference ``src/algorithms/emd_unifrac.py:58-61``) and its
    values — entries that fp32 operands would round away included."""
    from fununifrac_b200 import compute_fununifrac, sparse_npz
    from fununifrac_b200.utility import constant
    from oracle import oracle_c
    path, inp, ctx, leaves = full_tree
    n, N = ctx.n, 9
    names = [inp.idx_to_node[k] for k in range(n)]
    rng = np.random.default_rng(3)
    prof = tmp_path / "profiles"
    prof.mkdir()
    base = None
    for s in range(N):
        indptr, indices, values = synth.synth_profile_rows(leaves, n, [s], seed=17)
        ab = rng.geometric(0.2, len(indices)).astype(np.float64)
        if s == 0:
            base = (indices, ab)
        if s in (4, 5):                       # near-duplicates of sample 0: one abundance differs in the 9th digit
            indices, ab = base[0], base[1].copy()
            ab[s] *= 1 + 1e-9
        if s == 6:                            # identical to sample 0
            indices, ab = base
        synth.write_gather_csv(str(prof / f"s{s:02d}_gather.csv"), [names[j] for j in indices], ab)
    out = str(tmp_path / "out")
    compute_fununifrac.main(["-e", path, "-fd", str(prof), "-o", out, "-i", "d", "-b", "ko00001", "-a", "median_abund",
                             "-t", "4", "--diffab"] + (["--L2"] if l2 else []))
    files = sorted(glob.glob(str(prof / "*_gather.csv")))
    _, X = make_emd_input.native_profiles_to_matrix(files, inp, "median_abund", False, 4)
    P = oracle_c.push_up(inp.parent, inp.edge_length, X, l2)
    want = oracle_c.diffab_dense(P, l2)                                  # [N, N, n], antisymmetrised like the reference
    coo = sparse_npz.load_npz(os.path.join(out, constant.FUNUNIFRAC_OUT__DIFFAB_FILE_NAME.format("d")) + ".npz")
    assert coo.shape == (N, N, n)
    got = coo.todense()
    scale = np.abs(want).max()
    # the production float64 profiles differ from the reference's by at most a few ulp of the entry on the ~2 % of nodes
    # whose subtree crosses a 32-node chunk (summed chunk-wise instead of child by child), exact elsewhere: a difference
    # vector entry carries that noise in absolute terms
    pmax = np.abs(P).max() ** (2 if l2 else 1)
    noise = 16 * np.finfo(np.float64).eps * pmax
    assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want) + noise)
    # non-zero pattern: identical except where the entry itself is below that noise
    diff_pat = (got != 0) != (want != 0)
    assert diff_pat.mean() < 1e-3 and (not diff_pat.any() or np.maximum(np.abs(want), np.abs(got))[diff_pat].max() <= noise)
    assert not got[0, 6].any() and not got[6, 0].any()                   # identical samples: no entries at all
    assert got[0, 4].any() and 0 < np.abs(got[0, 4]).max() < 1e-6 * scale    # near-duplicates: tiny, but present
    # fp32 operands would have lost them: the largest entry of the near-duplicate row is below one fp32 ulp of the profiles
    assert np.abs(want[0, 4]).max() < np.finfo(np.float32).eps * np.abs(P[0]).max() ** (2 if l2 else 1)
    # --validate: float64 kernels in the reference's operation order -> values AND pattern bit-identical to the oracle
    compute_fununifrac.main(["-e", path, "-fd", str(prof), "-o", out, "-i", "v", "-b", "ko00001", "-a", "median_abund",
                             "-t", "4", "--diffab", "--validate"] + (["--L2"] if l2 else []))
    coo_v = sparse_npz.load_npz(os.path.join(out, constant.FUNUNIFRAC_OUT__DIFFAB_FILE_NAME.format("v")) + ".npz")
    assert np.array_equal(coo_v.todense(), want)
    D = np.load(os.path.join(out, constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("d")))
    Dref = oracle_c.pairwise(P, l2)
    iu = np.triu_indices(N, 1)
    ok = Dref[iu] > 0
    # near-duplicate pairs sit at the float64 noise floor of the push-up itself (the reference's own value there is the
    # rounding of its additions): agreement to a few ulp of the profile mass, 1e-6 relative everywhere else
    mass = (P ** 2).sum(1).max() if l2 else np.abs(P).sum(1).max()
    assert np.all(np.abs(D[iu][ok] - Dref[iu][ok]) <= REL * Dref[iu][ok] + 64 * np.finfo(np.float64).eps * mass)
    assert D[0, 6] == 0


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp_dir:
        result = run_config2(int(sys.argv[1]) if len(sys.argv) > 1 else 1000, tmp_dir)
    assert result["max_rel_err_vs_oracle"] < REL, result
    print(json.dumps(result))
