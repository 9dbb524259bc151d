"""Graph queries next to the hot path, answered from the device plan (GNNGraphs/src/query.jl):

    adjacency_list(g[, nodes]; dir=:out, with_eid=false)   query.jl:176-206   (the plan's CSR is that list)
    inneighbors / outneighbors(g, i)                        query.jl:109-141
    adjacency_matrix(g; dir=:out, weighted=true)           query.jl:220-231   (dense, for small graphs and tests)
    has_self_loops / has_multi_edges / is_bidirected        query.jl:553-579   (pair sort / duplicate runs on the device)
"""
from __future__ import annotations

from typing import List

import torch

from .graph import GNNGraph, _as_index
from .sampling import sample_edge_ids
from .transform import remove_multi_edges, sort_edge_index


def adjacency_list(g: GNNGraph, nodes=None, *, dir: str = "out", with_eid: bool = False):
    """Per queried node (default: all), its out-neighbours (dir="out": targets of its out-edges) or in-neighbours, in COO
    order; with_eid also returns the 1-based edge ids.  Lists of Python lists, like the reference's Vector{Vector}."""
    assert dir in ("out", "in")
    nodes = torch.arange(1, g.num_nodes + 1) if nodes is None else _as_index(nodes).reshape(-1)
    eids, offsets = sample_edge_ids(g, nodes, -1, dir=dir)          # K = -1: every incident edge, adjacency order
    dev = eids.device
    other = (g.t if dir == "out" else g.s).to(dev)[eids - 1].tolist()
    off = offsets.tolist()
    adj = [other[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    if not with_eid:
        return adj
    el = eids.tolist()
    return adj, [el[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def outneighbors(g: GNNGraph, i: int) -> List[int]:
    return adjacency_list(g, [i], dir="out")[0]


def inneighbors(g: GNNGraph, i: int) -> List[int]:
    return adjacency_list(g, [i], dir="in")[0]


def adjacency_matrix(g: GNNGraph, *, dir: str = "out", weighted: bool = True) -> torch.Tensor:
    """Dense A with A[i, j] = (summed weight of the) edges i -> j for dir="out", its transpose for dir="in"."""
    assert dir in ("out", "in")
    w = g.w if (weighted and g.w is not None) else None
    A = torch.zeros(g.num_nodes, g.num_nodes, dtype=torch.float32 if w is not None else torch.int64, device=g.s.device)
    vals = w if w is not None else torch.ones(g.num_edges, dtype=torch.int64, device=g.s.device)
    A.index_put_((g.s.long() - 1, g.t.long() - 1), vals, accumulate=True)
    return A if dir == "out" else A.t()


def has_self_loops(g: GNNGraph) -> bool:
    return bool((g.s == g.t).any())


def has_multi_edges(g: GNNGraph) -> bool:
    """more edges than distinct (s, t) pairs (query.jl:575-579)"""
    if g.num_edges == 0:
        return False
    plain = GNNGraph(g.s, g.t, num_nodes=g.num_nodes)
    return remove_multi_edges(plain).num_edges < g.num_edges


def is_bidirected(g: GNNGraph) -> bool:
    """sort_edge_index(s, t) == sort_edge_index(t, s) (query.jl:553-558)"""
    s1, t1 = sort_edge_index(g.s, g.t)
    s2, t2 = sort_edge_index(g.t, g.s)
    return bool(torch.equal(s1, s2) and torch.equal(t1, t2))
