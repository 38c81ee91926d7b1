"""numpy / pure-Python restatement of the reference arithmetic for SMALL cases.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: pinned in tests/test_oracle.py
against the reference's own outputs (tests/golden/*.npz).

Each function follows the cited reference lines literally (dict-free, same operation order).
"""
import itertools
import sys

import numpy as np

epsilon = sys.float_info.epsilon  # src/algorithms/emd_unifrac.py:10


def push_up(parent, length, P, is_l2=False):
    """src/algorithms/emd_unifrac.py:20-28 (L2) / :31-39 (L1), one vector, Python loop."""
    P_pushed = np.array(P, dtype=np.float64, copy=True)
    n = len(P_pushed)
    for i in range(n - 1):
        l = length[i]
        if l == 0:
            l = epsilon
        P_pushed[parent[i]] += P_pushed[i]
        P_pushed[i] *= np.sqrt(l) if is_l2 else l
    return P_pushed


def get_L1(P, Q):
    return np.sum(np.abs(P - Q))  # src/objects/profile_vector.py:9-11


def get_L2(P, Q):
    return np.sum(np.abs(P - Q) ** 2)  # :17-19


def pairwise(Ps, is_l2=False):
    """src/algorithms/emd_unifrac.py:42-54."""
    N = len(Ps)
    dists = np.zeros((N, N))
    for i, j in itertools.combinations(range(N), 2):
        dists[i, j] = dists[j, i] = get_L2(Ps[i], Ps[j]) if is_l2 else get_L1(Ps[i], Ps[j])
    return dists


def diffab_dense(Ps, is_l2=False):
    """:56-72 — [i,j,:] = P_i-P_j (or its square) for i<j, [j,i,:] = -[i,j,:]."""
    N, n = len(Ps), len(Ps[0])
    out = np.zeros((N, N, n))
    for i, j in itertools.combinations(range(N), 2):
        d = (Ps[i] - Ps[j]) ** 2 if is_l2 else Ps[i] - Ps[j]
        out[i, j] = d
        out[j, i] = -d
    return out


def diffab_coo(Ps, is_l2=False):
    """The COO the reference builds (:56-72), as sorted (coords[3, nnz], data[nnz]) of COO − COOᵀ."""
    dense = diffab_dense(Ps, is_l2)
    coords = np.array(np.nonzero(dense))
    return coords, dense[tuple(coords)]
