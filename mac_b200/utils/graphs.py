"""Graph containers and Laplacian builders (host side of the drop-in; mac/utils/graphs.py).

`Edge`, `select_edges` and the two CSR builders keep the reference names and argument
meaning (graphs.py:11, :13-48, :58-98, :101-111).  The builders return scipy CSR for API
compatibility only -- the accelerated path never materialises a host CSR; it assembles L(w)
on the device (`macb_set_x`).  The per-edge Python loops of the reference are replaced by
array construction ("next" row 2 of SURVEY section 8f).
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
from scipy.sparse import coo_matrix, csr_matrix

Edge = namedtuple("Edge", ["i", "j", "weight"])


def edges_to_arrays(edges):
    """list[Edge] (or an (i, j, w) array triple) -> (int32[E], int32[E], float64[E])."""
    # array triple only when the three members are arrays / plain lists -- a tuple of exactly three Edge tuples is an
    # edge list like any other iterable of Edge (the reference accepts any iterable)
    if (isinstance(edges, tuple) and len(edges) == 3 and not isinstance(edges, Edge)
            and all(isinstance(a, (np.ndarray, list)) for a in edges)):
        i, j, w = edges
        return np.asarray(i, dtype=np.int32), np.asarray(j, dtype=np.int32), np.asarray(w, dtype=np.float64)
    if len(edges) == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64)
    arr = np.asarray([(e[0], e[1], e[2]) for e in edges], dtype=np.float64)
    return arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32), arr[:, 2].copy()


def _laplacian(ei, ej, w, num_nodes):
    ei = np.asarray(ei, dtype=np.int64)
    ej = np.asarray(ej, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    rows = np.stack([ei, ej, ei, ej], axis=1).ravel()
    cols = np.stack([ei, ej, ej, ei], axis=1).ravel()
    data = np.stack([w, w, -w, -w], axis=1).ravel()
    return csr_matrix(coo_matrix((data, (rows, cols)), shape=[num_nodes, num_nodes]))


def weight_graph_lap_from_edge_list(edges, num_nodes: int) -> csr_matrix:
    """graphs.py:13-48."""
    ei, ej, w = edges_to_arrays(edges)
    return _laplacian(ei, ej, w, num_nodes)


def weight_graph_lap_from_edges(edges, weights, num_nodes: int) -> csr_matrix:
    """graphs.py:58-98: `edges` is int[E, 2], `weights` float[E]."""
    assert len(edges) == len(weights)
    edges = np.asarray(edges).reshape(-1, 2)
    return _laplacian(edges[:, 0], edges[:, 1], weights, num_nodes)


def weight_reduced_graph_lap_from_edge_list(edges, num_nodes: int) -> csr_matrix:
    """graphs.py:51-55."""
    return weight_graph_lap_from_edge_list(edges, num_nodes)[1:, 1:]


def select_edges(edges, w):
    """graphs.py:101-111."""
    assert len(edges) == len(w), f"Selection mask length {len(w)} does not match number of edges {len(edges)}"
    return [edge for i, edge in enumerate(edges) if w[i] == 1.0]


def get_edge_selection_as_binary_mask(edges, selected_edges) -> np.ndarray:
    """graphs.py:159-179."""
    assert len(edges) >= len(selected_edges), \
        "The number of selected edges cannot be greater than the total number of edges."
    chosen = set(selected_edges)
    return np.array([1.0 if e in chosen else 0.0 for e in edges])
