"""Tree → postorder integer arrays, gather CSV → abundance vector.

Mirrors ``src/factory/make_emd_input.py`` of the reference:
``tree_to_EMDU_input`` (:10-54), ``get_func_profiles_parallel`` (:58-66),
``parallel_functional_profile_to_vector`` (:69-79),
``functional_profile_to_vector`` (:82-122) and ``extend_vector`` (:124-138).

The arithmetic that defines parity with the reference — ``P / P.sum()`` on a
dense float64 vector with numpy's own summation — is done with the same numpy
calls so that the normalised vector is bit-identical; only the per-row Python
loop (``iterrows``) is replaced by vectorised parsing, with a row-by-row slow
path that keeps the reference's error order for malformed inputs.
"""
from __future__ import annotations

import multiprocessing
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from fununifrac_b200.objects.func_tree import FuncTree, FuncTreeEmduInput
from fununifrac_b200.objects.profile_vector import FuncProfileVector


def tree_to_EMDU_input(tree: FuncTree, brite=None, edge_len_property=None) -> FuncTreeEmduInput:
    """
    Convert a weighted graph to the format required by EMDUniFrac: Tint, lint, and nodes_in_order.
    Since EMDUniFrac wants everything to be integers, also return a mapping from integers to node names.

    Postorder contract: DFS from the unique in-degree-0 node, children in order
    of first appearance in the edge file, root last (``:42-43``).
    """
    if brite is not None:
        G = tree.set_subtree(brite)
    else:
        G = tree.main_tree
    if edge_len_property is None:
        edge_len_property = FuncTree.infer_edge_len_property(G)
    Tint = {}
    for node in G.nodes():
        pred = list(G.predecessors(node))
        if len(pred) > 1:
            raise Exception(f'Node {node} has more than one parent: {pred}')
        elif len(pred) == 1:
            Tint[node] = pred[0]
    lint = {}
    for i, j, data in G.edges(data=True):
        weight = data[edge_len_property]
        lint[(i, j)] = weight
        lint[(j, i)] = weight
    root = FuncTree.get_root_of_tree(G)
    nodes_in_order = G.dfs_postorder_nodes(root)
    node_2_EMDU_index = {node: i for i, node in enumerate(nodes_in_order)}
    Tint = {node_2_EMDU_index[node]: node_2_EMDU_index[Tint[node]] for node in Tint}
    lint = {(node_2_EMDU_index[i], node_2_EMDU_index[j]): lint[(i, j)] for i, j in lint}
    n = len(nodes_in_order)
    basis = list(range(n))
    EMDU_index_2_node = {i: node for node, i in node_2_EMDU_index.items()}
    parent = np.full(n, -1, dtype=np.int32)
    length = np.zeros(n, dtype=np.float64)
    for i, p in Tint.items():
        parent[i] = p
        length[i] = lint[(i, p)]
    out = FuncTreeEmduInput(Tint, lint, basis, EMDU_index_2_node, parent=parent, edge_length=length)
    out.__dict__["_node_to_idx"] = node_2_EMDU_index
    return out


# ---------------------------------------------------------------------------
# gather CSV → vector
# ---------------------------------------------------------------------------

def _ko_of(name: str) -> str:
    """``name`` column → node id (``:105-109``)."""
    if name.startswith('ko:'):
        return name.split(':')[-1]  # ko:K12345 case
    # abc|xyz|lmnop|ko:K12345 case
    return name.split('|')[-1].split(':')[-1]


def _scatter_slow(names, values, node_2_idx, n, abundance_key) -> np.ndarray:
    """Row-by-row restatement of the reference loop (error order preserved)."""
    P = np.zeros(n)
    for name, raw in zip(names, values):
        try:
            abundance = float(raw)
        except ValueError:
            raise ValueError(
                f"The abundance key {abundance_key} gives a value that is not a number: {raw}")
        try:
            ko = _ko_of(name)
            if ko not in node_2_idx:
                print(f"Warning: {ko} not found in EMDU index, skipping.")
            else:
                P[node_2_idx[ko]] = abundance
        except Exception:
            raise Exception(f"Could not parse the name {name} in the functional profile")
    return P


def scatter_profile(names: Sequence, values, node_2_idx: dict, n: int,
                    abundance_key: str = 'median_abund') -> np.ndarray:
    """Dense un-normalised float64[n] from (name, abundance) rows; duplicate node → last row wins."""
    vals = np.asarray(values)
    clean = vals.dtype.kind in "fiub" and all(type(x) is str for x in names)
    if not clean:
        return _scatter_slow(names, values, node_2_idx, n, abundance_key)
    P = np.zeros(n)
    vals = vals.astype(np.float64, copy=False)
    idx = np.empty(len(vals), dtype=np.int64)
    keep = np.ones(len(vals), dtype=bool)
    for r, name in enumerate(names):
        ko = _ko_of(name)
        j = node_2_idx.get(ko)
        if j is None:
            print(f"Warning: {ko} not found in EMDU index, skipping.")
            keep[r] = False
            idx[r] = 0
        else:
            idx[r] = j
    if not keep.all():
        idx, vals = idx[keep], vals[keep]
    # numpy assigns repeated indices in order, so the last row wins like the reference's loop
    P[idx] = vals
    return P


def functional_profile_to_vector(functional_profile, input: FuncTreeEmduInput, abundance_key='median_abund',
                                 normalize=True) -> FuncProfileVector:
    """This function will take a sourmash functional profile and convert it to a form that can be used by EMDUniFrac
    :param functional_profile: pandas DataFrame of the csv file output from `sourmash gather`
    :param input: FuncTreeEmduInput (for the node → index map)
    :param abundance_key: key in the functional profile that contains the abundance information
    :param normalize: whether to normalize the abundance information
    :return: vector P
    """
    df = functional_profile
    node_2_idx = input.node_to_idx()
    n = len(input.idx_to_node)
    if len(df) == 0:
        P = np.zeros(n)
    else:
        values = df[abundance_key]
        names = df['name']
        P = scatter_profile(list(names), values.to_numpy() if hasattr(values, "to_numpy") else values,
                            node_2_idx, n, abundance_key)
    if P.sum() == 0:
        raise Exception(f"Functional profile is empty! Perhaps try a different abundance key? "
                        f"You used {abundance_key}.")
    if normalize:
        P = P / P.sum()
    return P


def extend_vector(raw_P, input: FuncTreeEmduInput, normalize=True):
    """pandas Series indexed by KO → dense vector (``:124-138``; combined-dataframe caller)."""
    node_2_idx = input.node_to_idx()
    if normalize:
        raw_P = raw_P / raw_P.sum()
    P = np.zeros(len(input.idx_to_node))
    for ko in raw_P.index:
        if ko not in node_2_idx:
            print(f"Warning: {ko} not found in EMDU index, skipping.")
        else:
            P[node_2_idx[ko]] = raw_P[ko]
    return P


# ---------------------------------------------------------------------------
# process pool over files
# ---------------------------------------------------------------------------
_WORKER_INPUT: Optional[FuncTreeEmduInput] = None


def _worker_init(idx_to_node):
    global _WORKER_INPUT
    _WORKER_INPUT = FuncTreeEmduInput({}, {}, list(range(len(idx_to_node))), idx_to_node)


def _read_one(file, input, abundance_key, unweighted):
    import pandas as pd

    df = pd.read_csv(file)
    P = functional_profile_to_vector(df, input, abundance_key=abundance_key, normalize=True)
    if unweighted:
        # if entries are positive, set to 1 (after normalisation, not renormalised: :62-64)
        P[P > 0] = 1
    return file, P


def _read_one_worker(args):
    file, abundance_key, unweighted = args
    return _read_one(file, _WORKER_INPUT, abundance_key, unweighted)


def get_func_profiles_parallel(args):
    """(file, input, abundance_key, unweighted) → (file, P); same tuple protocol as the reference (:58-66)."""
    return _read_one(*args)


def parallel_functional_profile_to_vector(num_threads, fun_files, input: FuncTreeEmduInput, abundance_key,
                                          unweighted) -> List[Tuple[str, FuncProfileVector]]:
    """Pool(num_threads) over files, results in ``fun_files`` order (``:69-79``).

    The tree input is shipped to each worker once (pool initializer) instead of
    being pickled with every task.
    """
    fun_files = list(fun_files)
    if num_threads is None or num_threads <= 1 or len(fun_files) <= 1:
        return [_read_one(f, input, abundance_key, unweighted) for f in fun_files]
    # spawn, not fork: the parent may already hold a CUDA context
    ctx = multiprocessing.get_context("spawn")
    with ctx.Pool(num_threads, initializer=_worker_init, initargs=(input.idx_to_node,)) as pool:
        results = pool.map(_read_one_worker,
                           [(f, abundance_key, unweighted) for f in fun_files],
                           chunksize=max(2, len(fun_files) // num_threads))
    return results


def native_profiles_to_matrix(fun_files, input: FuncTreeEmduInput, abundance_key, unweighted, num_threads=0,
                              normalize=True, out: Optional[np.ndarray] = None) -> Tuple[List[str], np.ndarray]:
    """All gather CSVs → dense X[N, n] through the native ingest (``fuf_ingest_gather_csv``, csrc/ingest.cu): a C++
    thread pool that tokenises only the ``name`` and abundance columns of each file.  Row f is bit-for-bit what
    ``functional_profile_to_vector(pd.read_csv(fun_files[f]), input, abundance_key)`` returns (pinned on 800+ number
    formats, quoting, duplicates, NA); a file the native parser reports as malformed is re-read through that very
    pandas route, so odd inputs behave — return or raise — exactly as in the reference."""
    import ctypes

    from fununifrac_b200 import engine

    L = engine.load_library()
    fun_files = list(fun_files)
    n = len(input.idx_to_node)
    names = [str(input.idx_to_node[k]).encode() for k in range(n)]
    arr = (ctypes.c_char_p * n)(*names)
    ix = ctypes.c_void_p()
    engine._check(L.fuf_name_index_create(ctypes.byref(ix), arr, n), "fuf_name_index_create")
    try:
        N = len(fun_files)
        X = out if out is not None else np.zeros((N, n), dtype=np.float64)
        if X.dtype != np.float64 or X.shape[0] < N or X.shape[1] < n or X.strides[1] != 8:
            raise ValueError("out must be a C-ordered float64 [>=N, >=n] array")
        paths = (ctypes.c_char_p * max(N, 1))(*[os.fsencode(f) for f in fun_files])
        status = np.zeros(max(N, 1), dtype=np.int32)
        stride = 512
        messages = ctypes.create_string_buffer(stride * max(N, 1))
        wcap = 1 << 20
        warnings = ctypes.create_string_buffer(wcap)
        rc = L.fuf_ingest_gather_csv(ix, paths, N, str(abundance_key).encode(), int(bool(normalize)),
                                     int(bool(unweighted)), int(num_threads or 0), X.ctypes.data,
                                     X.strides[0] // 8 if N else n, status.ctypes.data, messages, stride, warnings, wcap)
        if rc < 0:
            engine._check(rc, "fuf_ingest_gather_csv")
        # Inputs the native tokeniser rejects or reads differently from pandas (bool / '1_0' / overflowing abundances,
        # rows with an extra field = pandas' implicit index column, numeric names, an all-NaN profile ...) are rare and
        # cheap to hand to the reference's own route: whatever pandas + functional_profile_to_vector return or raise
        # for that file is then, by construction, the reference's behaviour.
        redo = {f for f in range(N) if status[f] in (1, 2, 3, 4)}
        for line in warnings.value.decode("utf-8", "replace").splitlines():
            fidx, ko = line.split(chr(9), 1)
            if int(fidx) not in redo:
                print(f"Warning: {ko} not found in EMDU index, skipping.")
        for f in sorted(redo):
            import pandas as pd
            P = functional_profile_to_vector(pd.read_csv(fun_files[f]), input, abundance_key=abundance_key,
                                             normalize=normalize)
            if unweighted:
                P[P > 0] = 1
            X[f, :n] = P
            status[f] = 0
        for f in range(N):
            if status[f]:
                msg = messages.raw[f * stride:(f + 1) * stride].split(b"\0", 1)[0].decode("utf-8", "replace")
                if status[f] == 1:
                    raise ValueError(f"The abundance key {abundance_key} gives a value that is not a number: {msg}")
                if status[f] == 2:
                    raise Exception(f"Functional profile is empty! Perhaps try a different abundance key? "
                                    f"You used {abundance_key}.")
                if status[f] == 4:
                    raise Exception(f"Could not parse the name {msg} in the functional profile")
                if status[f] == 5:
                    raise KeyError(msg)
                raise Exception(f"{fun_files[f]}: {msg}")
        return fun_files, X[:N]
    finally:
        L.fuf_name_index_destroy(ix)


def profiles_to_matrix(results, n: int) -> Tuple[List[str], np.ndarray]:
    """Stack ``(file, P)`` pairs into the dense sample-major float64 matrix X[N, n] the C-ABI takes."""
    files = []
    rows = []
    for file, P in results:
        files.append(file)
        rows.append(P)
    X = np.empty((len(rows), n), dtype=np.float64)
    for i, P in enumerate(rows):
        X[i, :] = P
    return files, X
