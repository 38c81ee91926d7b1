"""Minimal 3-D COO container + ``.npz`` I/O in the layout pydata ``sparse`` uses.

The reference stores ``--diffab`` output with ``sparse.save_npz`` (``src/compute_all_pairwise_fununifrac.py:94``)
and its tests read it back with ``sparse.load_npz(...)[i, j, :].todense()``
(``tests/scripts/test_make_all_pw_fununifrac.py:78-88``).  ``sparse`` is an unpinned third-party package that
is not installed here; its published on-disk format is a (compressed) ``.npz`` with four arrays
``data``, ``coords`` (ndim x nnz, lexicographically sorted), ``shape`` and ``fill_value`` — restated here.
Byte-level parity with the package is unpinned (no copy of it to test against); value-level parity is
pinned by the reference's own tests restated in ``tests/``.

If the real package is importable, ``to_pydata_sparse`` converts to a genuine ``sparse.COO``.
"""
from __future__ import annotations

import numpy as np


class COO:
    """Sorted, duplicate-free COO array (the subset of ``sparse.COO`` the path and its tests touch)."""

    def __init__(self, coords, data, shape, fill_value=0.0, sorted_=False):
        coords = np.asarray(coords, dtype=np.int64).reshape(len(shape), -1)
        data = np.asarray(data)
        if not sorted_ and coords.shape[1]:
            order = np.lexsort(coords[::-1])
            coords, data = coords[:, order], data[order]
        self.coords, self.data, self.shape = coords, data, tuple(int(s) for s in shape)
        self.fill_value = np.asarray(fill_value, dtype=data.dtype if data.size else np.float64)[()]

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nnz(self):
        return self.coords.shape[1]

    @property
    def dtype(self):
        return self.data.dtype

    def todense(self):
        out = np.full(self.shape, self.fill_value, dtype=self.data.dtype if self.data.size else np.float64)
        if self.nnz:
            out[tuple(self.coords)] = self.data
        return out

    def transpose(self, axes=None):
        axes = tuple(range(self.ndim))[::-1] if axes is None else tuple(axes)
        return COO(self.coords[list(axes)], self.data, [self.shape[a] for a in axes], self.fill_value)

    def __neg__(self):
        return COO(self.coords, -self.data, self.shape, -self.fill_value, sorted_=True)

    def __sub__(self, other):
        if self.shape != other.shape:
            raise ValueError("shape mismatch")
        coords = np.concatenate([self.coords, other.coords], axis=1)
        data = np.concatenate([self.data, -other.data])
        order = np.lexsort(coords[::-1])
        coords, data = coords[:, order], data[order]
        if coords.shape[1]:
            new = np.ones(coords.shape[1], dtype=bool)
            new[1:] = np.any(coords[:, 1:] != coords[:, :-1], axis=0)
            starts = np.nonzero(new)[0]
            data = np.add.reduceat(data, starts)
            coords = coords[:, starts]
            keep = data != 0
            coords, data = coords[:, keep], data[keep]
        return COO(coords, data, self.shape, self.fill_value - other.fill_value, sorted_=True)

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (self.ndim - len(key))
        mask = np.ones(self.nnz, dtype=bool)
        keep_axes, new_shape = [], []
        for ax, k in enumerate(key):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (self.shape[ax] if k < 0 else 0)
                if not 0 <= k < self.shape[ax]:
                    raise IndexError("index out of range")
                mask &= self.coords[ax] == k
            elif isinstance(k, slice) and k == slice(None):
                keep_axes.append(ax)
                new_shape.append(self.shape[ax])
            else:
                raise NotImplementedError("only integer indices and full slices are supported")
        if not keep_axes:
            return self.data[mask][0] if mask.any() else self.fill_value
        return COO(self.coords[keep_axes][:, mask], self.data[mask], new_shape, self.fill_value, sorted_=True)

    def take(self, *indices) -> np.ndarray:
        """Dense ``self.todense()[np.ix_(*indices)]`` without densifying the whole array (one integer index list per
        axis; negative and repeated indices allowed) — what ``DiffabArrayIndexer`` needs from a loaded diffab file."""
        if len(indices) != self.ndim:
            raise IndexError(f"take needs {self.ndim} index lists")
        uniq, inv, pos = [], [], []
        for ax, ind in enumerate(indices):
            ind = np.asarray(ind, dtype=np.int64).reshape(-1)
            size = self.shape[ax]
            if ind.size and (ind.min() < -size or ind.max() >= size):
                raise IndexError(f"index out of bounds for axis {ax} with size {size}")
            u, iv = np.unique(np.where(ind < 0, ind + size, ind), return_inverse=True)
            lut = np.full(size, -1, dtype=np.int64)
            lut[u] = np.arange(len(u))
            uniq.append(u)
            inv.append(iv.reshape(-1))
            pos.append(lut[self.coords[ax]])
        block = np.full([len(u) for u in uniq], self.fill_value, dtype=self.data.dtype if self.data.size else np.float64)
        if self.nnz:
            hit = np.all(np.stack(pos) >= 0, axis=0)
            block[tuple(p[hit] for p in pos)] = self.data[hit]
        return block[np.ix_(*inv)]


def save_npz(filename, matrix: COO, compressed=True):
    """Same arrays and names as ``sparse.save_npz`` (numpy appends ``.npz`` when missing)."""
    nodes = {"data": matrix.data, "coords": matrix.coords, "shape": np.asarray(matrix.shape),
             "fill_value": np.asarray(matrix.fill_value)}
    (np.savez_compressed if compressed else np.savez)(filename, **nodes)


def load_npz(filename) -> COO:
    with np.load(filename) as fp:
        return COO(fp["coords"], fp["data"], tuple(fp["shape"]), fp["fill_value"][()], sorted_=True)


def to_pydata_sparse(matrix: COO):
    import sparse  # optional

    return sparse.COO(matrix.coords, matrix.data, shape=matrix.shape, fill_value=matrix.fill_value)
