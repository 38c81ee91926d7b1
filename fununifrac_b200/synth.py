"""Deterministic synthetic KO-like trees and sample profiles.

The benchmark and the parity tests need inputs of the real KO tree's shape
without shipping KEGG data: a 6-level hierarchy ``root → ko00001 → 8 → 56 →
480 → 3,456 → 26,000 leaves`` (30,002 nodes / 30,001 edges), Zipf-skewed
fan-out, log-normal branch lengths with a few exact zeros and a few
``2.22e-16`` (the deterministic KEGG tree of the reference has 145 such edges),
written as the 3-column TSV the reference's ``import_graph`` reads unchanged.

Profiles: per sample a Zipf-popular subset of leaves (shared core) with
``1 + Geometric`` integer abundances, normalised to 1 in float64.  Every sample
is generated from ``default_rng([seed, sample_index])`` so any row block of any
N is reproducible on its own.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np

FULL_WIDTHS = (1, 1, 8, 56, 480, 3456, 26000)
EPS = 2.220446049250313e-16
GATHER_HEADER = ("intersect_bp,f_orig_query,f_match,f_unique_to_query,f_unique_weighted,average_abund,"
                 "median_abund,std_abund,filename,name,md5,f_match_orig,unique_intersect_bp,gather_result_rank,"
                 "remaining_bp,query_filename,query_name,query_md5,query_bp,ksize,moltype,scaled,query_n_hashes,"
                 "query_abundance,query_containment_ani,match_containment_ani,average_containment_ani,"
                 "max_containment_ani,potential_false_negative")


def _zipf_assign(rng, n_children: int, n_parents: int, a: float = 1.1) -> np.ndarray:
    """Parent slot of each child: every parent gets ≥1 child, the rest Zipf-skewed."""
    if n_parents >= n_children:
        return np.arange(n_children) % n_parents
    w = 1.0 / np.arange(1, n_parents + 1) ** a
    rng.shuffle(w)
    extra = rng.choice(n_parents, size=n_children - n_parents, p=w / w.sum())
    assign = np.concatenate([np.arange(n_parents), extra])
    assign.sort(kind="stable")
    return assign


def synth_ko_tree_edges(seed: int = 0, widths: Sequence[int] = FULL_WIDTHS
                        ) -> List[Tuple[str, str, float]]:
    """Edge list (parent, child, length) in depth-first file order."""
    rng = np.random.default_rng([seed, 0x7EE])
    depth = len(widths)
    names: List[List[str]] = []
    for lvl, wdt in enumerate(widths):
        if lvl == 0:
            names.append(["root"])
        elif lvl == 1 and wdt == 1:
            names.append(["ko00001"])
        elif lvl == depth - 1:
            names.append([f"K{j:05d}" for j in range(wdt)])
        else:
            names.append([f"L{lvl}_{j:05d} level {lvl} group" for j in range(wdt)])
    children: List[List[List[int]]] = []
    for lvl in range(depth - 1):
        assign = _zipf_assign(rng, widths[lvl + 1], widths[lvl])
        kids: List[List[int]] = [[] for _ in range(widths[lvl])]
        for c, p in enumerate(assign):
            kids[int(p)].append(c)
        children.append(kids)
    n_edges = sum(widths[1:])
    # log-normal fitted to the AAI tree: mean ≈ 0.11, clipped to [5e-5, 0.5]
    lengths = np.clip(rng.lognormal(mean=-2.6, sigma=0.9, size=n_edges), 5e-5, 0.5)
    special = rng.random(n_edges)
    lengths[special < 0.005] = 0.0
    lengths[(special >= 0.005) & (special < 0.010)] = EPS
    edges: List[Tuple[str, str, float]] = []
    e = 0
    stack = [(0, 0, 0)]  # (level, index, next child position)
    while stack:
        lvl, idx, pos = stack.pop()
        if lvl == depth - 1:
            continue
        kids = children[lvl][idx]
        if pos < len(kids):
            stack.append((lvl, idx, pos + 1))
            c = kids[pos]
            edges.append((names[lvl][idx], names[lvl + 1][c], float(lengths[e])))
            e += 1
            stack.append((lvl + 1, c, 0))
    return edges


def write_edge_list(path: str, edges, header: str = "#parent\tchild\tedge_length") -> str:
    with open(path, "w") as f:
        f.write(header + "\n")
        for p, c, l in edges:
            f.write(f"{p}\t{c}\t{l!r}\n")
    return path


def synth_ko_tree(path: str, seed: int = 0, widths: Sequence[int] = FULL_WIDTHS) -> str:
    """Write the synthetic tree as a TSV edge list and return ``path``."""
    return write_edge_list(path, synth_ko_tree_edges(seed, widths))


def synth_profile_rows(leaf_index: np.ndarray, n: int, rows: Sequence[int], seed: int = 1,
                       nnz_lo: float = 2000 / 26000, nnz_hi: float = 8000 / 26000,
                       zipf_a: float = 0.9):
    """CSR pieces for the given sample indices.

    :param leaf_index: postorder indices of the leaves (the columns that can be non-zero)
    :param n: number of tree nodes (vector length)
    :return: (indptr int64[len(rows)+1], indices int32[nnz], values float64[nnz]) — values normalised per row
    """
    leaf_index = np.asarray(leaf_index)
    n_leaf = len(leaf_index)
    pop_rng = np.random.default_rng([seed, 0xA11])
    logp = -zipf_a * np.log(np.arange(1, n_leaf + 1, dtype=np.float64))
    logp = logp[pop_rng.permutation(n_leaf)]
    lo = max(1, int(round(nnz_lo * n_leaf)))
    hi = max(lo, int(round(nnz_hi * n_leaf)))
    indptr = [0]
    idx_chunks, val_chunks = [], []
    for r in rows:
        rng = np.random.default_rng([seed, int(r)])
        k = int(rng.integers(lo, hi + 1))
        keys = logp + rng.gumbel(size=n_leaf)
        chosen = np.argpartition(keys, n_leaf - k)[n_leaf - k:]
        chosen.sort()
        abund = 1.0 + rng.geometric(0.15, size=k).astype(np.float64)
        vec = np.zeros(n)
        vec[leaf_index[chosen]] = abund
        vec = vec / vec.sum()  # same float64 normalisation as functional_profile_to_vector
        cols = leaf_index[chosen].astype(np.int32)
        order = np.argsort(cols, kind="stable")
        idx_chunks.append(cols[order])
        val_chunks.append(vec[cols[order]])
        indptr.append(indptr[-1] + k)
    return (np.asarray(indptr, dtype=np.int64),
            np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, np.int32),
            np.concatenate(val_chunks) if val_chunks else np.zeros(0, np.float64))


def synth_profiles_dense(leaf_index: np.ndarray, n: int, N: int, seed: int = 1, dtype=np.float64,
                         out: np.ndarray = None, row0: int = 0, **kw) -> np.ndarray:
    """Dense sample-major X[N, n] (rows ``row0 .. row0+N-1`` of the seeded set)."""
    if out is None:
        out = np.zeros((N, n), dtype=dtype)
    else:
        out[:] = 0
    block = 256
    for b in range(0, N, block):
        rows = range(row0 + b, row0 + min(N, b + block))
        indptr, indices, values = synth_profile_rows(leaf_index, n, rows, seed=seed, **kw)
        for i in range(len(rows)):
            s, e = indptr[i], indptr[i + 1]
            out[b + i, indices[s:e]] = values[s:e]
    return out


def write_gather_csv(path: str, names: Sequence[str], abundances: Sequence[float],
                     key: str = "median_abund") -> str:
    """A gather CSV with the 29 sourmash columns of ``data/small_sim_10_KOs_gather.csv:1``;
    only ``name`` and the abundance columns carry information."""
    cols = GATHER_HEADER.split(",")
    with open(path, "w") as f:
        f.write(GATHER_HEADER + "\n")
        for rank, (nm, ab) in enumerate(zip(names, abundances)):
            row = {c: "0" for c in cols}
            row.update(name=f"ko:{nm}", filename="synthetic.sig.zip", md5="0" * 32, query_filename="synth.fa",
                       query_name="", query_md5="00000000", moltype="DNA", query_abundance="TRUE",
                       potential_false_negative="FALSE", gather_result_rank=str(rank), ksize="7", scaled="1")
            for c in ("median_abund", "average_abund", "f_unique_weighted", "f_orig_query", "f_match",
                      "f_unique_to_query"):
                row[c] = repr(float(ab))
            row[key] = repr(float(ab))
            f.write(",".join(row[c] for c in cols) + "\n")
    return path


def leaves_of(parent: np.ndarray) -> np.ndarray:
    """Postorder indices of leaves given the parent array."""
    n = len(parent)
    has_child = np.zeros(n, dtype=bool)
    has_child[parent[parent >= 0]] = True
    return np.nonzero(~has_child)[0]
