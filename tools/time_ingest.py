"""Ingest timing: Python mirror of the reference route (pandas + pool) vs the native C++ ingest, same files."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import synth
from fununifrac_b200.factory import make_emd_input, make_tree

if __name__ == "__main__":
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    with tempfile.TemporaryDirectory() as tmp:
        path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
        inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
        n = len(inp.basis); leaves = synth.leaves_of(inp.parent); names = [inp.idx_to_node[k] for k in range(n)]
        files = []
        for i in range(nf):
            indptr, indices, values = synth.synth_profile_rows(leaves, n, [i], seed=1)
            ab = 1.0 + np.random.default_rng(i).geometric(0.15, len(indices))
            files.append(synth.write_gather_csv(os.path.join(tmp, f"s{i:04d}_gather.csv"), [names[j] for j in indices], ab))
        print("rows in last file", len(indices), "bytes", os.path.getsize(files[-1]))
        t = time.time(); res = make_emd_input.parallel_functional_profile_to_vector(1, files, inp, "median_abund", False)
        Xp = make_emd_input.profiles_to_matrix(res, n)[1]; tp = time.time() - t
        t = time.time(); _, Xn = make_emd_input.native_profiles_to_matrix(files, inp, "median_abund", False, 1); tn1 = time.time() - t
        t = time.time(); _, Xn = make_emd_input.native_profiles_to_matrix(files, inp, "median_abund", False, 8); tn8 = time.time() - t
        print(f"{nf} files: python mirror 1 process {tp:.2f} s ({tp / nf * 1e3:.1f} ms/file); native 1 thread {tn1:.3f} s "
              f"({tn1 / nf * 1e3:.2f} ms/file), 8 threads {tn8:.3f} s; bit-identical {np.array_equal(Xp, Xn)}")
