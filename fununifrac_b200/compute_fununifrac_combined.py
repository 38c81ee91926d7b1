#!/usr/bin/env python
"""All-pairs FunUniFrac from ONE combined KO x sample table — same flags, outputs and quirks as the reference's
``fununifrac/compute_fununifrac_combined.py:30-74``, computed by the CUDA engine.

    python -m fununifrac_b200.compute_fununifrac_combined -e tree.txt -f table.csv -o out/ -b ko00001

Kept from the reference:
* the table is read with ``pd.read_csv(file, index_col='name')``; columns are samples, the output order is the column
  order (written to ``{out_id}.basis.npy`` as text);
* every column is normalised by its FULL sum (KOs that are not in the tree count in the sum, are then skipped with a
  warning, ``make_fununifrac_inputs`` :13-27);
* profiles are always pushed up with the L1 weights, also with ``--L2`` (:60-63), and ``--L2`` then only selects the
  squared pairwise form; ``-a``, ``--unweighted`` and ``--Ppushed`` are accepted and ignored (:41-53);
* outputs: ``{out_dir}/{out_id}.npy`` (float64 N x N) and ``{out_dir}/{out_id}.basis.npy``; ``out_id`` defaults to
  ``fununifrac_out``; ``--diffab`` computes the difference vectors but (like the reference) does not save them.
"""
import argparse
import logging

import numpy as np

import fununifrac_b200.factory.make_emd_input as make_emd_input
import fununifrac_b200.factory.make_tree as make_tree
from fununifrac_b200 import cli_flags, engine
from fununifrac_b200.algorithms.emd_unifrac import EarthMoverDistanceUniFracSolver
from fununifrac_b200.objects.func_tree import FuncTreeEmduInput


def make_fununifrac_inputs(raw_P, input: FuncTreeEmduInput, normalize=True):
    """pandas Series indexed by KO -> dense vector over the tree basis (reference :13-27)."""
    return make_emd_input.extend_vector(raw_P, input, normalize)


def table_to_matrix(sample_df, input: FuncTreeEmduInput, normalize=True) -> np.ndarray:
    """All columns of the table at once: X[N, n] with the arithmetic of ``make_fununifrac_inputs`` per column:
    ``column / column.sum()`` — ``Series.sum()``: NaN skipped, numpy's pairwise summation over the contiguous column —
    scattered to the basis; rows whose KO is not in the tree are dropped with one warning each.  A table with a repeated
    KO id goes column by column through ``make_fununifrac_inputs`` itself, so that it fails the way the reference
    does (``P[i] = raw_P[ko]`` with a two-element selection)."""
    node_2_idx = input.node_to_idx()
    n = len(input.idx_to_node)
    if not sample_df.index.is_unique:
        return np.stack([make_fununifrac_inputs(sample_df[c], input, normalize) for c in sample_df.columns])
    vals_t = np.ascontiguousarray(sample_df.to_numpy(dtype=np.float64).T)        # [N, rows]: one contiguous row per sample
    if normalize:
        sums = np.where(np.isnan(vals_t), 0.0, vals_t).sum(axis=1)             # pairwise per row, as Series.sum()
        vals_t = vals_t / sums[:, None]
    X = np.zeros((vals_t.shape[0], n), dtype=np.float64)
    for r, ko in enumerate(sample_df.index):
        j = node_2_idx.get(ko)
        if j is None:
            print(f"Warning: {ko} not found in EMDU index, skipping.")
        else:
            X[:, j] = vals_t[:, r]
    return X


def build_parser() -> argparse.ArgumentParser:
    return cli_flags.build(
        "All-pairs functional UniFrac between the sample columns of one combined KO x sample table, over a KEGG "
        "Orthology tree with edge lengths, computed on a B200.",
        ["edge_list", "file", "out_dir", "out_id", "abundance_key", "brite", "diffab", "verbose", "unweighted", "L2",
         "Ppushed", "validate"])


def compute(sample_df, input: FuncTreeEmduInput, is_L2: bool, make_diffab: bool = False, validate: bool = False):
    """The compute section (:57-67) as one device job: L1 push-up of every column, then all pairs in the requested
    form.  Returns (dists, diffabs or None)."""
    import torch

    ctx = engine.context_for(input)
    X = table_to_matrix(sample_df, input)
    N = X.shape[0]
    mode = engine.MODE_L2 if is_L2 else engine.MODE_L1
    Xd = torch.from_numpy(X).to(f"cuda:{ctx.device}")
    if validate:
        P64T = ctx.push_up_f64(Xd, engine.MODE_L1)
        dists = ctx.pairwise_f64(P64T, N, mode).cpu().numpy()
        PT = P64T
    else:
        pushed = ctx.push_up_job(Xd, engine.MODE_L1)                 # always the L1 weights (reference quirk)
        job = ctx.round_pushed(pushed.P64T, N, mode)                 # rounding statistics in the pairwise form's norm
        dists = ctx.distances(job, mode).cpu().numpy()
        PT = pushed.P64T                                             # difference vectors in float64, like P - Q
    diffabs = EarthMoverDistanceUniFracSolver(validate=validate).diffab_coo(ctx, PT, N, mode) if make_diffab else None
    return dists, diffabs


def main(argv=None):
    import pandas as pd

    args = build_parser().parse_args(argv)
    logging.basicConfig(level=args.loglevel, format='%(asctime)s %(levelname)s: %(message)s')
    tree = make_tree.import_graph(args.edge_list, directed=True)
    input: FuncTreeEmduInput = make_emd_input.tree_to_EMDU_input(tree, args.brite)
    sample_df = pd.read_csv(args.file, index_col='name')
    dists, _ = compute(sample_df, input, args.L2, args.diffab, args.validate)
    out_id = args.out_id if args.out_id else "fununifrac_out"
    out_file = f"{args.out_dir}/{out_id}.npy"
    np.save(out_file, dists)
    basis_out = f"{args.out_dir}/{out_id}.basis.npy"
    with open(basis_out, 'w') as f:
        for file in sample_df.columns:
            f.write(f"{file}\n")


if __name__ == '__main__':
    main()
