#!/usr/bin/env python
"""BASELINE config 2 through the drop-in CLI: synthetic full KO tree (30,002 nodes), N gather CSVs on disk
(29 sourmash columns, 2,000-8,000 KO rows each), weighted L1, median_abund, one B200.

    python tests/config2_cli.py [N=1000]

Times the whole user-visible run (argument parsing -> tree -> native ingest -> push-up + all pairs -> np.save) and
checks the saved matrix against the CPU oracle on the same files.  Prints one JSON line."""
import glob
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fununifrac_b200 import synth  # noqa: E402
from fununifrac_b200.factory import make_emd_input, make_tree  # noqa: E402


def _write(args):
    out, leaves, n, names, i = args
    indptr, indices, values = synth.synth_profile_rows(leaves, n, [i], seed=1)
    ab = 1.0 + np.random.default_rng([7, i]).geometric(0.15, len(indices))
    synth.write_gather_csv(os.path.join(out, f"s{i:05d}_gather.csv"), [names[j] for j in indices], ab)
    return len(indices)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    from fununifrac_b200 import compute_fununifrac
    from fununifrac_b200.utility import constant
    from oracle import oracle_c
    with tempfile.TemporaryDirectory() as tmp:
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
        nbytes = sum(os.path.getsize(f) for f in glob.glob(os.path.join(prof, "*_gather.csv")))
        argv = ["-e", tree, "-fd", prof, "-o", os.path.join(tmp, "out"), "-i", "cfg2", "-b", "ko00001", "-a", "median_abund",
                "-t", str(os.cpu_count() or 1)]
        compute_fununifrac.main(argv + ["--force"])                     # first call: CUDA context, library load
        t = time.perf_counter()
        compute_fununifrac.main(argv + ["--force"])
        t_cli = time.perf_counter() - t
        D = np.load(os.path.join(tmp, "out", constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format("cfg2")))
        # oracle on the same files
        files = sorted(glob.glob(os.path.join(prof, "*_gather.csv")))
        t = time.perf_counter()
        _, X = make_emd_input.native_profiles_to_matrix(files, inp, "median_abund", False, os.cpu_count() or 1)
        t_ingest = time.perf_counter() - t
        oracle_c.set_threads(os.cpu_count() or 1)
        t = time.perf_counter()
        Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, False), False)
        t_cpu = time.perf_counter() - t
        iu = np.triu_indices(N, 1)
        err = float((np.abs(D[iu] - Dref[iu]) / Dref[iu]).max())
        assert err < 1e-6, err
        print(json.dumps({
            "config": f"BASELINE config 2: {N} gather CSVs ({rows} KO rows, {nbytes / 1e6:.0f} MB), synthetic full KO tree "
                      f"(n={n}), weighted L1, median_abund, CLI on one GPU",
            "cli_seconds": t_cli, "of_which_native_ingest_seconds": t_ingest, "pairs": N * (N - 1) // 2,
            "pairs_per_second_whole_cli": N * (N - 1) // 2 / t_cli, "max_rel_err_vs_oracle": err,
            "cpu_oracle_port_seconds_compute_only": t_cpu, "threads": os.cpu_count(), "csv_generation_seconds": t_gen}))


if __name__ == "__main__":
    main()
