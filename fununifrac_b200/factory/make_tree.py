"""Edge-list TSV → ``FuncTree``.

Restates the observable behaviour of the reference's ``import_graph``
(``src/factory/make_tree.py:6-43``), including its header sniffing:

1. first attempt = ``nx.read_edgelist(file, delimiter='\\t', nodetype=str)``
   (``:19``): ``#`` starts a comment anywhere in a line, lines with fewer than
   two tab fields are skipped, and any third column must be a Python dict
   literal — a plain number there makes the attempt fail;
2. fallback (``:24-37``): line 1 is consumed as a header and must split into
   exactly three tab fields, the third naming the edge attribute; the remaining
   lines are parsed with ``data=((col3, float),)`` so every edge line must have
   exactly three fields and a float-parsable length.

Any failure of the fallback surfaces as the reference's generic ``Exception``
(``:38-41``); a missing file is a ``ValueError`` (``:15-16``).
"""
from __future__ import annotations

import os
from ast import literal_eval

from fununifrac_b200.objects.func_tree import FuncTree, TreeDiGraph

_COMMENT = "#"


def _decoded_lines(fid, encoding="utf-8"):
    for line in fid:
        yield line if isinstance(line, str) else line.decode(encoding)


def _strip_comment(line: str):
    p = line.find(_COMMENT)
    if p >= 0:
        line = line[:p]
    return line


def _parse_edgelist(lines, data) -> TreeDiGraph:
    """``networkx.parse_edgelist`` semantics for ``delimiter='\\t'``.

    ``data=True`` → a third column is evaluated as a dict literal;
    ``data=((key, type),)`` → typed columns.
    """
    G = TreeDiGraph()
    for line in lines:
        line = _strip_comment(line)
        if not line:
            continue
        s = line.rstrip("\n").split("\t")
        if len(s) < 2:
            continue
        u = s.pop(0)
        v = s.pop(0)
        d = s
        if len(d) == 0 or data is False:
            edgedata = {}
        elif data is True:
            try:
                edgedata = dict(literal_eval(" ".join(d).strip()))
            except Exception as err:
                raise TypeError(f"Failed to convert edge data ({d}) to dictionary.") from err
        else:
            if len(d) != len(data):
                raise IndexError(f"Edge data {d} and data_keys {data} are not the same length")
            edgedata = {}
            for (edge_key, edge_type), edge_value in zip(data, d):
                try:
                    edge_value = edge_type(edge_value)
                except Exception as err:
                    raise TypeError(
                        f"Failed to convert {edge_key} data {edge_value} to type {edge_type}.") from err
                edgedata[edge_key] = edge_value
        G.add_edge(u, v, **edgedata)
    return G


def import_graph(edge_list_file, directed=True) -> FuncTree:
    """
    Import a graph from an edge list. The edge list can have 2 or 3 columns corresponding to parent, child,
    and (optionally) edge length.

    :param edge_list_file: text file with edge list, tab separated
    :param directed: must be True (the all-pairs path only ever passes True,
        ``src/compute_all_pairwise_fununifrac.py:53``)
    :return: FuncTree
    """
    if not os.path.exists(edge_list_file):
        raise ValueError(f"Could not find edge list file {edge_list_file}")
    if not directed:
        raise NotImplementedError("only directed trees are on the all-pairs FunUniFrac path")
    try:
        with open(edge_list_file, 'rb') as fid:
            G = _parse_edgelist(_decoded_lines(fid), data=True)
    except Exception:
        try:
            with open(edge_list_file, 'rb') as fid:
                line = fid.readline()
                try:
                    line = line.decode('utf-8')
                    col1, col2, col3 = line.strip().split('\t')
                except Exception:
                    raise ValueError('The first line of the file must have at most three columns: parent child '
                                     'edge_length')
                G = _parse_edgelist(_decoded_lines(fid), data=((col3, float),))
        except Exception:
            raise Exception(
                'Could not read edge list file. Make sure it is a tab-delimited file with two columns: parent & child'
                'or three columns: parent, child, edge_length')
    return FuncTree(G)
