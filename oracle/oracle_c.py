"""ctypes binding of ``oracle/oracle.c`` (TEST INFRASTRUCTURE — see the header of oracle.c).

Parity status: pinned against the reference's own outputs in ``tests/test_oracle.py``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i32p, f64p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
        L.oracle_push_up.argtypes = [i32p, f64p, ctypes.c_int32, f64p, f64p, ctypes.c_int]
        L.oracle_push_up_batch.argtypes = [i32p, f64p, ctypes.c_int32, f64p, f64p, ctypes.c_int64, ctypes.c_int]
        L.oracle_pairwise.argtypes = [f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, f64p]
        L.oracle_pairwise_rows.argtypes = [f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, ctypes.c_int64,
                                           ctypes.c_int64, f64p]
        L.oracle_diffab_dense.argtypes = [f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int, f64p]
        L.oracle_solve.argtypes = [i32p, f64p, ctypes.c_int32, f64p, f64p, f64p]
        L.oracle_solve.restype = ctypes.c_double
        L.oracle_np_sum.argtypes = [f64p, ctypes.c_int64]
        L.oracle_np_sum.restype = ctypes.c_double
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def push_up(parent, length, X, is_l2=False):
    """X: float64[N, n] (or [n]) un-pushed profiles → pushed profiles, same shape."""
    X = np.asarray(X, dtype=np.float64)
    single = X.ndim == 1
    X2 = np.ascontiguousarray(X.reshape(1, -1) if single else X)
    N, n = X2.shape
    parent, pp = _i32(parent)
    length, lp = _f64(length)
    X2, xp = _f64(X2)
    out = np.empty_like(X2)
    lib().oracle_push_up_batch(pp, lp, n, xp, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), N, int(is_l2))
    return out[0] if single else out


def pairwise(P, is_l2=False):
    """P: float64[N, n] pushed profiles → D float64[N, N] (reference pairwise_computation without diffab)."""
    P, pp = _f64(P)
    N, n = P.shape
    D = np.empty((N, N), dtype=np.float64)
    lib().oracle_pairwise(pp, N, n, int(is_l2), D.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return D


def pairwise_rows(P, r0, r1, is_l2=False):
    P, pp = _f64(P)
    N, n = P.shape
    D = np.empty((r1 - r0, N), dtype=np.float64)
    lib().oracle_pairwise_rows(pp, N, n, int(is_l2), r0, r1, D.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return D


def diffab_dense(P, is_l2=False):
    P, pp = _f64(P)
    N, n = P.shape
    out = np.empty((N, N, n), dtype=np.float64)
    lib().oracle_diffab_dense(pp, N, n, int(is_l2), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def solve(parent, length, P, Q):
    parent, pp = _i32(parent)
    length, lp = _f64(length)
    P, p = _f64(P)
    Q, q = _f64(Q)
    n = len(P)
    diffab = np.empty(n, dtype=np.float64)
    Z = lib().oracle_solve(pp, lp, n, p, q, diffab.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return Z, diffab


def np_sum(a):
    a, p = _f64(a)
    return lib().oracle_np_sum(p, a.size)


def num_threads():
    return lib().oracle_num_threads()


def set_threads(n):
    """Use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1, which would otherwise cap the baseline)."""
    lib().oracle_set_threads(int(n))
