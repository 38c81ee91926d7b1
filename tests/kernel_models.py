"""numpy models of the CUDA kernels' index math and summation order (TEST CODE).

They restate, step by step, what ``csrc/pushup.cu`` and ``csrc/pairwise.cu`` do on the device so that the
host-side planning (``fuf_debug_tree_plan``) and the numerical scheme (fp32 partial sums over 256 basis
entries folded into double) can be checked against the oracle on a box without a GPU.  The GPU tests check
the real kernels against the same oracle.
"""
import ctypes

import numpy as np

from fununifrac_b200 import engine


def tree_plan(parent):
    L = engine.load_library()
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    n = len(parent)
    pl, cp = np.empty(n, np.int32), np.empty(n, np.int32)
    sp, span = np.empty(n, np.uint8), np.empty(5 * n, np.int32)
    ns, ncp, contig, chunk = (ctypes.c_int32() for _ in range(4))
    L.fuf_debug_tree_plan.restype = ctypes.c_int
    rc = L.fuf_debug_tree_plan(parent.ctypes.data_as(ctypes.c_void_p), n, pl.ctypes.data_as(ctypes.c_void_p),
                               cp.ctypes.data_as(ctypes.c_void_p), sp.ctypes.data_as(ctypes.c_void_p),
                               span.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ns), ctypes.byref(ncp),
                               ctypes.byref(contig), ctypes.byref(chunk))
    if rc != 0:
        raise RuntimeError(L.fuf_last_error().decode())
    return dict(parent_local=pl, cp_after=cp, is_span=sp.astype(bool), span=span[:5 * ns.value].reshape(-1, 5),
                n_cp=ncp.value, contiguous=bool(contig.value), chunk=chunk.value)


def pushup_model(parent, length, X, is_l2, plan=None, x_dtype=np.float64):
    """K1a/K1b/K1c of pushup.cu on the host: returns PT-like float32 [N, n]."""
    parent = np.asarray(parent)
    n = len(parent)
    plan = plan or tree_plan(parent)
    assert plan["contiguous"]
    KT = plan["chunk"]
    w = np.array(length, dtype=np.float64)
    w[:-1][w[:-1] == 0] = np.finfo(np.float64).eps
    if is_l2:
        w = np.sqrt(w)
    w[-1] = 1.0
    X = np.asarray(X).astype(x_dtype)
    N = X.shape[0]
    n_chunks = (n + KT - 1) // KT
    out = np.full((N, n), np.nan, dtype=np.float32)
    T = np.zeros((n_chunks, N))
    CP = np.zeros((max(1, plan["n_cp"]), N))
    for c in range(n_chunks):                       # K1a: one CTA per chunk (and 128-sample block)
        k0 = c * KT
        kt = min(KT, n - k0)
        acc = np.zeros((KT, N))
        r = np.zeros(N)
        for kk in range(kt):
            k = k0 + kk
            x = X[:, k].astype(np.float64)
            S = x + acc[kk]
            r = r + x
            pl = plan["parent_local"][k]
            if pl >= 0:
                acc[pl] += S
            if not plan["is_span"][k]:
                out[:, k] = (w[k] * S).astype(np.float32)
            cp = plan["cp_after"][k]
            if cp >= 0:
                CP[cp] = r
        T[c] = r
    CT = np.zeros_like(T)                            # K1b: exclusive scan over chunks
    run = np.zeros(N)
    for c in range(n_chunks):
        CT[c] = run
        run = run + T[c]
    for k, cf, cl, s_start, s_k in plan["span"]:     # K1c: spanning nodes from prefix differences
        hi = CT[cl] + CP[s_k]
        lo = CT[cf] + (CP[s_start] if s_start >= 0 else 0.0)
        out[:, k] = (w[k] * (hi - lo)).astype(np.float32)
    assert not np.isnan(out).any()
    return out


def push_meta(parent):
    """Per-node plan words of the TMA push-up kernel (``fuf_debug_push_meta``), or None when it does not apply."""
    L = engine.load_library()
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    n = len(parent)
    meta = np.empty((n + 31) // 32 * 32, np.int32)
    usable = ctypes.c_int32()
    L.fuf_debug_push_meta.restype = ctypes.c_int
    rc = L.fuf_debug_push_meta(parent.ctypes.data_as(ctypes.c_void_p), n, meta.ctypes.data_as(ctypes.c_void_p),
                               ctypes.byref(usable))
    if rc != 0:
        raise RuntimeError(L.fuf_last_error().decode())
    return meta if usable.value else None


def pushup_tma_model(parent, length, X, is_l2, center=None, plan=None):
    """K1t (pushup_tma_kernel) + K1b on the host: thread = sample; per chunk a running prefix r, S = x for leaf-like
    nodes, S = r - parked prefix otherwise, spanning nodes from chunk totals and checkpoints.  Returns (float32 [N, n],
    per-sample rounding statistic)."""
    parent = np.asarray(parent)
    n = len(parent)
    plan = plan or tree_plan(parent)
    meta = push_meta(parent)
    assert meta is not None
    KT = 32
    w = np.array(length, dtype=np.float64)
    w[:-1][w[:-1] == 0] = np.finfo(np.float64).eps
    if is_l2:
        w = np.sqrt(w)
    w[-1] = 1.0
    c = np.zeros(n) if center is None else np.asarray(center, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    n_chunks = (n + KT - 1) // KT
    out = np.full((N, n), np.nan, dtype=np.float32)
    err = np.zeros(N)
    T = np.zeros((n_chunks, N))
    CP = np.zeros((max(1, plan["n_cp"]), N))

    def emit(k, v):
        f = v.astype(np.float32)
        out[:, k] = f
        e = v - f.astype(np.float64)
        return e * e if is_l2 else np.abs(e)

    for ch in range(n_chunks):
        k0 = ch * KT
        r = np.zeros(N)
        parked = {}
        for kk in range(KT):
            k = k0 + kk
            m = int(meta[k])
            x = X[:, k] if k < n else np.zeros(N)
            r = r + x
            keep, take = (m >> 4) & 15, (m >> 8) & 15
            if keep:
                parked[keep - 1] = r.copy()
            if not (m & 2):
                if m & 1:
                    v = w[k] * x
                else:
                    S = r - parked[take - 1] if take else r
                    v = w[k] * S - c[k]
                err += emit(k, v)
            cp = (m >> 16) & 0xFFFF
            assert bool(cp) == bool(m & 8)
            if cp:
                CP[cp - 1] = r
        T[ch] = r
    for k, cf, cl, s_start, s_k in plan["span"]:     # K1b, same-sign form
        head = T[cf] - (CP[s_start] if s_start >= 0 else 0.0)
        mid = T[cf + 1:cl].sum(0) if cl > cf + 1 else 0.0
        err += emit(k, w[k] * ((head + CP[s_k]) + mid) - c[k])
    assert not np.isnan(out).any()
    return out, err


def pairwise_model(P32, is_l2, flush=256):
    """K2's numerical scheme: fp32 subtraction and accumulation over chunks of `flush` basis entries,
    chunk sums combined in double.  P32: float32 [N, n]."""
    P32 = np.asarray(P32, dtype=np.float32)
    N, n = P32.shape
    D = np.zeros((N, N))
    for k0 in range(0, n, flush):
        blk = P32[:, k0:k0 + flush]
        d = blk[:, None, :] - blk[None, :, :]
        d = np.abs(d) if not is_l2 else d * d
        part = np.zeros((N, N), dtype=np.float32)
        for kk in range(d.shape[2]):                 # sequential fp32 adds, like one register accumulator
            part = part + d[:, :, kk]
        D += part.astype(np.float64)
    np.fill_diagonal(D, 0.0)
    return D
