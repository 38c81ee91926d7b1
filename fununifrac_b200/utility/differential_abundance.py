"""Consumers of ``--diffab`` output: mirror of ``src/utility/differential_abundance.py:7-84``.

``DiffabArrayIndexer`` and ``convert_diffab_array_to_df`` keep the reference's names, arguments, return shapes and
exceptions.  The reference slices a materialised N x N x n array (dense, or the ``sparse.COO`` written by
``sparse.save_npz``).  Here the array may be any of

* a dense ``numpy.ndarray`` (reference behaviour, ``np.ix_`` slicing),
* the :class:`fununifrac_b200.sparse_npz.COO` that ``sparse_npz.load_npz`` returns for a ``*.diffab.npz`` file,
* a :class:`PushedDiffabStore`: the GPU-resident store.  It keeps only the pushed matrix (N x n, node-major float64 in
  HBM) and evaluates the requested ``[files1, files2, nodes]`` block on demand with ``fuf_diffab_gather``
  (csrc/diffab.cu), because the tensor itself is ``T[i, j, :] = P_i - P_j`` (L1) or ``sign(j - i) (P_i - P_j)^2`` (L2):
  N x N x n never has to exist.  At N = 10,000, n = 30,002 that is 2.4 GB resident instead of 24 TB.

The plotting helpers of the reference module (matplotlib figures) are out of scope (not on the path).
"""
from __future__ import annotations

import os.path
from typing import Dict, Mapping, Optional, Sequence

import numpy as np

from .. import engine
from ..sparse_npz import COO


class PushedDiffabStore:
    """GPU-resident diffab tensor, stored as the pushed profiles it is a function of.

    ``store[i, j]`` / ``store[i, j, :]`` / ``store[i, j, k]`` index like the dense array; :meth:`gather` returns the
    outer-product block the indexer needs.  Values are float64 differences of the float64 pushed vectors: identical to
    ``get_L1_diffab`` / ``get_L2_diffab`` (``src/objects/profile_vector.py:13-23``) on the same pushed vectors.
    """

    def __init__(self, ctx: "engine.TreeContext", P64T, N: int, is_L2: bool):
        self.ctx, self.P64T, self.N = ctx, P64T, int(N)
        self.mode = engine.MODE_L2 if is_L2 else engine.MODE_L1
        self.shape = (self.N, self.N, ctx.n)
        self.ndim = 3
        self.dtype = np.dtype(np.float64)

    @classmethod
    def from_pushed(cls, Ps_pushed: Mapping[str, np.ndarray], file_basis: Sequence[str], input, is_L2: bool,
                    device: Optional[int] = None) -> "PushedDiffabStore":
        """Build from the dict the driver pickles with ``--Ppushed`` (``src/compute_all_pairwise_fununifrac.py:99-102``)."""
        torch = engine._torch()
        ctx = engine.context_for(input, device)
        names = list(file_basis)
        N, n = len(names), ctx.n
        P64T = ctx.new_profile_buffer(max(N, 1), dtype=torch.float64)
        block = max(1, min(max(N, 1), (256 << 20) // (8 * n)))
        for b in range(0, N, block):
            rows = np.stack([np.asarray(Ps_pushed[nm], dtype=np.float64) for nm in names[b:b + block]])
            if rows.shape[1] != n:
                raise ValueError(f"pushed vectors must have length {n}")
            ctx.load_pushed(torch.from_numpy(rows).to(f"cuda:{ctx.device}"), P64T, col0=b)
        return cls(ctx, P64T, N, is_L2)

    def gather(self, rows_i, rows_j, nodes=None) -> np.ndarray:
        """``out[a, b, c] = T[rows_i[a], rows_j[b], nodes[c]]`` as a host float64 array (nodes=None: all, in order)."""
        out = self.ctx.diffab_gather(self.P64T, self.N, self.mode, rows_i, rows_j, nodes)
        return out.cpu().numpy()

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) > 3:
            raise IndexError("too many indices")
        key = key + (slice(None),) * (3 - len(key))
        idx, squeeze = [], []
        for ax, k in enumerate(key):
            size = self.shape[ax]
            if isinstance(k, (int, np.integer)):
                if not -size <= int(k) < size:
                    raise IndexError(f"index {k} is out of bounds for axis {ax} with size {size}")
                idx.append(np.array([int(k) % size]))
                squeeze.append(ax)
            elif isinstance(k, slice):
                idx.append(np.arange(size)[k])
            else:
                idx.append(np.asarray(k, dtype=np.int64).reshape(-1))
        out = self.gather(idx[0], idx[1], None if (isinstance(key[2], slice) and key[2] == slice(None)) else idx[2])
        return out.squeeze(axis=tuple(squeeze)) if squeeze else out

    def todense(self) -> np.ndarray:
        """The whole N x N x n array — small inputs only, like the reference's dense ``diffabs``."""
        return self.gather(np.arange(self.N), np.arange(self.N))


def _take(diffabs, indices1, indices2, indices3) -> np.ndarray:
    """``diffabs[np.ix_(indices1, indices2, indices3)]`` for the three kinds of store."""
    if isinstance(diffabs, PushedDiffabStore):
        return diffabs.gather(indices1, indices2, indices3)
    if isinstance(diffabs, COO):
        return diffabs.take(indices1, indices2, indices3)
    return diffabs[np.ix_(indices1, indices2, indices3)]


class DiffabArrayIndexer:
    """Index the differential-abundance array by file names and node names (reference :7-55)."""

    def __init__(self, diffabs, nodes_in_order, file_basis, EMDU_index_2_node: Dict[int, str]):
        self.diffabs = diffabs
        self.node_to_index = {node: i for i, node in enumerate(nodes_in_order)}
        self.file_to_index = {file: i for i, file in enumerate(file_basis)}
        self.node_to_EMDU_index = {v: k for k, v in EMDU_index_2_node.items()}

    def get_diffab(self, files1, files2) -> np.ndarray:
        """Differential-abundance vectors between two sets of files (sets can be single file names):
        array ``[len(files1), len(files2), n]`` (reference :18-33)."""
        if isinstance(files1, str):
            files1 = [files1]
        if isinstance(files2, str):
            files2 = [files2]
        indices1 = np.array([self.file_to_index[file] for file in files1], dtype=np.int64)
        indices2 = np.array([self.file_to_index[file] for file in files2], dtype=np.int64)
        return _take(self.diffabs, indices1, indices2, np.arange(self.diffabs.shape[2]))

    def get_diffab_for_node(self, files1, files2, nodes) -> np.ndarray:
        """Same, restricted to node name(s), e.g. ``'K01234'``: ``[len(files1), len(files2), len(nodes)]``
        (reference :35-55)."""
        if isinstance(files1, str):
            files1 = [files1]
        if isinstance(files2, str):
            files2 = [files2]
        if isinstance(nodes, (str, int)):
            nodes = [nodes]
        indices1 = np.array([self.file_to_index[file] for file in files1], dtype=np.int64)
        indices2 = np.array([self.file_to_index[file] for file in files2], dtype=np.int64)
        indices3 = np.array([self.node_to_index[self.node_to_EMDU_index[node]] for node in nodes], dtype=np.int64)
        return _take(self.diffabs, indices1, indices2, indices3)


def convert_diffab_array_to_df(diffabs, nodes_in_order, file_basis):
    """3-level table ``df[file][file2][node] = diffabs[i, j, k]`` (reference :58-84; columns = ``file``, index =
    ``file2``, each cell a dict over nodes).  Tiny inputs only: N*N*n Python objects."""
    import pandas as pd

    file_basis = [os.path.basename(x) for x in file_basis]
    num_samples = diffabs.shape[0]
    if num_samples != len(file_basis):
        raise ValueError('Number of samples does not match number of file names')
    if num_samples != diffabs.shape[1]:
        raise ValueError('Differential abundance array is not square')
    if diffabs.shape[2] != len(nodes_in_order):
        raise ValueError('Differential abundance array does not match nodes in order')
    dense = diffabs.todense() if hasattr(diffabs, "todense") else np.asarray(diffabs)
    nodes = list(nodes_in_order)
    table = {file: {file2: dict(zip(nodes, dense[i, j])) for j, file2 in enumerate(file_basis)}
             for i, file in enumerate(file_basis)}
    return pd.DataFrame.from_dict(table)
