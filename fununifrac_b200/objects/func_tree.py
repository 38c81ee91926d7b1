"""Tree objects of the all-pairs FunUniFrac path.

Mirrors the public surface of the reference's ``src/objects/func_tree.py``
(``FuncTreeEmduInput`` :5-10, ``FuncTree`` :13-138) without depending on
networkx: the hot path only needs insertion-ordered adjacency, a descendant
set, a unique root and a DFS postorder, all of which are provided by the small
``TreeDiGraph`` below with the same observable ordering rules as the
networkx ``DiGraph`` the reference uses (children are iterated in order of the
first appearance of each ``(parent, child)`` edge; re-adding an edge only
updates its attributes).

On top of the four dict/list fields the reference exposes, ``FuncTreeEmduInput``
carries the dense postorder arrays (``parent``, ``edge_length``) that the C-ABI
consumes (``include/fununifrac.h``: ``fuf_tree_create``).
"""
from __future__ import annotations

import dataclasses
import sys
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np

EPSILON = sys.float_info.epsilon  # reference: src/algorithms/emd_unifrac.py:10


class GraphError(Exception):
    """Raised where networkx would raise ``NetworkXError`` (e.g. unknown node)."""


class TreeDiGraph:
    """Insertion-ordered directed graph: the subset of ``networkx.DiGraph`` the path uses.

    Ordering contract (what makes the postorder bit-exact with the reference):
    nodes are ordered by first appearance (parent before child within one edge),
    and the successors of a node by the first appearance of the edge.
    """

    def __init__(self) -> None:
        self._succ: Dict[str, Dict[str, dict]] = {}
        self._pred: Dict[str, Dict[str, dict]] = {}

    # -- construction -----------------------------------------------------
    def add_node(self, u) -> None:
        if u not in self._succ:
            self._succ[u] = {}
            self._pred[u] = {}

    def add_edge(self, u, v, **attr) -> None:
        self.add_node(u)
        self.add_node(v)
        datadict = self._succ[u].get(v, {})
        datadict.update(attr)
        self._succ[u][v] = datadict
        self._pred[v][u] = datadict

    # -- queries ----------------------------------------------------------
    def __contains__(self, n) -> bool:
        try:
            return n in self._succ
        except TypeError:
            return False

    def __len__(self) -> int:
        return len(self._succ)

    def __iter__(self) -> Iterator:
        return iter(self._succ)

    def nodes(self) -> List:
        return list(self._succ)

    def number_of_nodes(self) -> int:
        return len(self._succ)

    def number_of_edges(self) -> int:
        return sum(len(nbrs) for nbrs in self._succ.values())

    def successors(self, n) -> Iterator:
        try:
            return iter(self._succ[n])
        except KeyError as err:
            raise GraphError(f"The node {n} is not in the digraph.") from err

    neighbors = successors

    def predecessors(self, n) -> Iterator:
        try:
            return iter(self._pred[n])
        except KeyError as err:
            raise GraphError(f"The node {n} is not in the digraph.") from err

    def in_degree(self, n=None):
        if n is not None:
            return len(self._pred[n])
        return ((node, len(p)) for node, p in self._pred.items())

    def out_degree(self, n=None):
        if n is not None:
            return len(self._succ[n])
        return ((node, len(s)) for node, s in self._succ.items())

    def edges(self, data: bool = False):
        for u, nbrs in self._succ.items():
            for v, d in nbrs.items():
                yield (u, v, d) if data else (u, v)

    def get_edge_data(self, u, v, default=None):
        try:
            return self._succ[u][v]
        except KeyError:
            return default

    def subgraph(self, nodes: Iterable) -> "TreeDiGraph":
        """Induced subgraph that preserves the parent graph's ordering.

        networkx returns a filtered *view*; the adjacency of a view iterates the
        underlying dicts (``networkx/classes/coreviews.py`` FilterAtlas), so the
        child order seen by ``dfs_postorder_nodes`` is the insertion order of the
        full graph.  A materialised copy built in that order is equivalent for
        everything the path observes.
        """
        keep = {n for n in nodes if n in self._succ}
        sub = TreeDiGraph()
        for n in self._succ:
            if n in keep:
                sub.add_node(n)
        for u, nbrs in self._succ.items():
            if u not in keep:
                continue
            for v, d in nbrs.items():
                if v in keep:
                    sub._succ[u][v] = d
                    sub._pred[v][u] = d
        return sub

    def descendants(self, source) -> set:
        """All nodes reachable from ``source`` (excluded), like ``nx.descendants``."""
        if source not in self._succ:
            raise GraphError(f"The node {source} is not in the graph.")
        seen = {source}
        stack = [source]
        while stack:
            u = stack.pop()
            for v in self._succ[u]:
                if v not in seen:
                    seen.add(v)
                    stack.append(v)
        seen.discard(source)
        return seen

    def dfs_postorder_nodes(self, source) -> List:
        """Iterative DFS postorder from ``source``, children in adjacency order.

        Same visiting rule as ``nx.dfs_postorder_nodes(G, source=root)`` used at
        ``src/factory/make_emd_input.py:43``: a child already visited is skipped.
        """
        if source not in self._succ:
            raise KeyError(source)
        order: List = []
        visited = {source}
        stack: List[Tuple[object, Iterator]] = [(source, iter(self._succ[source]))]
        while stack:
            parent, children = stack[-1]
            for child in children:
                if child not in visited:
                    visited.add(child)
                    stack.append((child, iter(self._succ[child])))
                    break
            else:
                stack.pop()
                order.append(parent)
        return order

    def to_networkx(self):
        """Materialise as ``networkx.DiGraph`` (compatibility helper, off the hot path)."""
        import networkx as nx  # optional dependency

        G = nx.DiGraph()
        for n in self._succ:
            G.add_node(n)
        for u, v, d in self.edges(data=True):
            G.add_edge(u, v, **d)
        return G


@dataclasses.dataclass
class FuncTreeEmduInput:
    """EMDUniFrac view of the tree; same four public fields as the reference
    (``src/objects/func_tree.py:5-10``) plus dense postorder arrays.

    ``parent[i]`` is the postorder index of the parent of node ``i`` (``-1`` for
    the root, which is always ``n-1``); ``edge_length[i]`` is the raw length of
    the edge ``(i, parent[i])`` (root slot 0.0).  Zero lengths are *not* yet
    replaced by epsilon here: the reference does that lazily inside ``push_up_*``
    (``src/algorithms/emd_unifrac.py:24-25,35-36``) and the solver mirrors it.
    """

    Tint: dict  # ancestor dictionary
    lint: dict  # lengths dictionary
    basis: list
    idx_to_node: dict  # emdu index to node
    parent: Optional[np.ndarray] = dataclasses.field(default=None, repr=False, compare=False)
    edge_length: Optional[np.ndarray] = dataclasses.field(default=None, repr=False, compare=False)

    def dense_arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        """(parent int32[n], edge_length float64[n]) rebuilt from the *current* dicts.

        ``lint`` is part of the public, mutable surface (the reference mutates
        it), so the arrays handed to the device always reflect it.
        """
        n = len(self.basis)
        parent = np.full(n, -1, dtype=np.int32)
        length = np.zeros(n, dtype=np.float64)
        Tint, lint = self.Tint, self.lint
        for i in range(n - 1):
            p = Tint[i]
            parent[i] = p
            length[i] = lint[i, p]
        return parent, length

    def node_to_idx(self) -> Dict[str, int]:
        cached = self.__dict__.get("_node_to_idx")
        if cached is None or len(cached) != len(self.idx_to_node):
            cached = {v: k for k, v in self.idx_to_node.items()}
            self.__dict__["_node_to_idx"] = cached
        return cached


class FuncTree:
    """Wrapper with brite-subtree selection (reference ``src/objects/func_tree.py:13-138``)."""

    def __init__(self, G: TreeDiGraph) -> None:
        self._G: TreeDiGraph = G
        self._subtree: Optional[TreeDiGraph] = None

    @property
    def main_tree(self) -> TreeDiGraph:
        return self._G

    @property
    def current_subtree(self) -> TreeDiGraph:
        self.check_subtree_valid()
        return self._subtree

    @property
    def basis(self):
        self.check_subtree_valid()
        return [x for x in self._subtree.nodes()]

    @property
    def basis_index(self):
        self.check_subtree_valid()
        return {node: i for i, node in enumerate(self.basis)}

    def check_subtree_valid(self):
        if self._subtree is None:
            raise Exception("subtree has not been made")

    def set_subtree(self, classification):
        """descendants(brite) ∪ {brite} ∪ {'root'} → induced subgraph (:44-53)."""
        G = self._G
        descendants = FuncTree.get_descendants(G, classification)
        descendants.add('root')
        subgraph = G.subgraph(descendants)
        self._subtree = subgraph
        return subgraph

    @classmethod
    def get_descendants(cls, G: TreeDiGraph, node, leaf_only=False) -> set:
        descendants = set()
        if leaf_only:
            for n in G.descendants(node):
                if G.out_degree(n) == 0:
                    descendants.add(n)
        else:
            descendants.add(node)
            descendants |= G.descendants(node)
        return descendants

    @classmethod
    def get_descendant(cls, G: TreeDiGraph, v1, v2):
        if v1 in G.predecessors(v2):
            return v2
        elif v2 in G.predecessors(v1):
            return v1
        else:
            raise ValueError(f"Nodes {v1} and {v2} are not adjacent")

    @classmethod
    def get_root_of_tree(cls, G: TreeDiGraph):
        roots = [n for n, d in G.in_degree() if d == 0]
        if len(roots) > 1:
            raise Exception(f"The graph has multiple roots: {roots}")
        return roots[0]

    @classmethod
    def infer_edge_len_property(cls, G: TreeDiGraph):
        first_edge = next(iter(G.edges(data=True)), None)      # the attribute names of the FIRST edge decide
        if first_edge is None:
            raise IndexError("list index out of range")        # what indexing an empty edge list raises
        edge_properties = list(first_edge[-1].keys())
        if len(edge_properties) > 1:
            raise ValueError(f'I found multiple edge properties: {edge_properties}. I don\'t know which one to use for '
                             f'edge lengths.')
        else:
            edge_len_property = edge_properties[0]
        return edge_len_property

    @classmethod
    def get_leaf_nodes(cls, G: TreeDiGraph):
        leaf_nodes = set()
        for n in G.nodes():
            if G.out_degree(n) == 0:
                leaf_nodes.add(n)
        return sorted(list(leaf_nodes))
