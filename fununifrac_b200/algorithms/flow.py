"""Host-side flow bookkeeping for the legacy ``solve_with_flow`` entry point
(reference ``src/algorithms/emd_unifrac.py:119-194``; "Not used. Moved to pushup paradigm").

EMDUniFrac flow: walking the tree in postorder, every node greedily matches the surplus sources
(P > Q) against the deficit sources (P < Q) that are still unmatched in its subtree; what is left is
carried to the parent with the source's distance grown by the edge length.  Sources that meet at a
node come from different child subtrees, so the node is their lowest common ancestor and the cost of a
matched unit is exactly the tree distance.

This is control-plane style bookkeeping on Python dicts (the flow F is a dict keyed by node pairs); it is
kept on the host.  The arithmetic that scales — subtree sums of P-Q — runs on the device (see
``EarthMoverDistanceUniFracSolver.solve_with_flow``).
"""
from __future__ import annotations

import numpy as np


def flow_matching(parent: np.ndarray, length: np.ndarray, P: np.ndarray, Q: np.ndarray):
    """Return ``(Z, F, reached)``.

    ``F[(j, k)]`` = mass moved from node j to node k (``F[(i, i)] = min(P[i], Q[i])`` where both are
    positive); ``Z`` = Σ tree-distance × mass; ``reached[i]`` is True when unmatched mass leaves node i
    towards its parent (the keys of the reference's diffab dict).  Sources are matched in ascending index
    order, the order CPython iterates small sets of small ints in, which is what the reference does.
    """
    n = len(P)
    F = dict()
    Z = 0
    w = np.zeros(n)
    pos = [[] for _ in range(n)]  # [source, remaining surplus]
    neg = [[] for _ in range(n)]  # [source, remaining deficit (negative)]
    reached = np.zeros(n, dtype=bool)
    for i in range(n):
        if P[i] > 0 and Q[i] > 0:
            F[(i, i)] = np.minimum(P[i], Q[i])
        g = P[i] - Q[i]
        if g > 0:
            pos[i].append([i, g])
        elif g < 0:
            neg[i].append([i, g])
        pl, nl = sorted(pos[i]), sorted(neg[i])
        a = b = 0
        while a < len(pl) and b < len(nl):
            j, k = pl[a][0], nl[b][0]
            val = np.minimum(pl[a][1], -nl[b][1])
            if val > 0:
                F[(j, k)] = val
                pl[a][1] = pl[a][1] - val
                nl[b][1] = nl[b][1] + val
                Z = Z + (w[j] + w[k]) * val
            step = False
            if pl[a][1] == 0:
                a += 1
                step = True
            if nl[b][1] == 0:
                b += 1
                step = True
            if not step:  # cannot happen for finite inputs; guards against an endless loop on NaN
                break
        rest_p, rest_n = pl[a:], nl[b:]
        pos[i], neg[i] = [], []
        if i < n - 1:
            par = int(parent[i])
            for src, _ in rest_p:
                w[src] += length[i]
            for src, _ in rest_n:
                w[src] += length[i]
            reached[i] = bool(rest_p or rest_n)
            pos[par].extend(rest_p)
            neg[par].extend(rest_n)
    return Z, F, reached
