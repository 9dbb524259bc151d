"""Neighbour sampling for mini-batch training over the device plan:

    sample_neighbors(g, nodes, K=-1; dir=:in, replace=false, dropnodes=false)   GNNGraphs/src/sampling.jl:68-121
    induced_subgraph(g, nodes)                                                  GNNGraphs/src/sampling.jl:172-204
    NeighborLoader(g; num_neighbors, input_nodes, num_layers, batch_size)       GNNGraphs/src/samplers.jl:28-105

The edge selection runs on the plan's CSR (csrc/sample.cu: one query touches only the rows asked for, where the
reference scans every edge through a Dict); what follows — slicing s, t, w and the edge features by the chosen edge
ids, relabelling nodes for `dropnodes` — is index bookkeeping on the device.

Random draws are counter based (`seed`); the reference draws from Julia's RNG, so agreement is distributional.
Two documented deviations: (1) a node listed twice in `nodes` is sampled twice (the reference's Dict keeps one entry);
(2) `NeighborLoader` expands one frontier per layer for the whole batch instead of one per input node (same union of
sampled neighbourhoods in distribution when frontiers do not overlap, fewer draws when they do).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from . import graph as _graph
from ._lib import lib
from .graph import GNNGraph, _as_index, _stream, rows, unrows

_seed_counter = [0x5EED]


def _next_seed(seed: Optional[int]) -> int:
    if seed is not None:
        return int(seed) & (2 ** 64 - 1)
    _seed_counter[0] = (_seed_counter[0] * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
    return _seed_counter[0]


def _take_last(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return unrows(rows(x)[idx.to(x.device)])


def sample_edge_ids(g: GNNGraph, nodes, K: int = -1, *, dir: str = "in", replace: bool = False,
                    seed: Optional[int] = None):
    """(eids, offsets): 1-based COO ids of the sampled edges, node after node, and the (len(nodes)+1) running counts."""
    assert dir in ("in", "out")
    p = g.plan()
    dev = p.device
    nodes = _as_index(nodes).to(dev).reshape(-1).contiguous()
    n = int(nodes.numel())
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    total = C.c_int64(0)
    d = _lib.DIR_IN if dir == "in" else _lib.DIR_OUT
    sd = _next_seed(seed)
    args = (p.h, nodes.data_ptr() if n else None, n, nodes.element_size(), 1, int(K), d, int(bool(replace)), sd)
    with torch.cuda.device(dev):
        _lib.check(lib.gnnb_sample_neighbors(*args, offsets.data_ptr(), None, 0, C.byref(total), _stream(dev)))
        eids = torch.empty(int(total.value), dtype=torch.int64, device=dev)
        if total.value:
            _lib.check(lib.gnnb_sample_neighbors(*args, offsets.data_ptr(), eids.data_ptr(), int(total.value),
                                                 C.byref(total), _stream(dev)))
    return eids, offsets


def _stable_unique(x: torch.Tensor) -> torch.Tensor:
    """distinct values of x in order of first appearance (Julia's `unique` / `setdiff` order)"""
    if x.numel() == 0:
        return x
    u, inv = torch.unique(x, return_inverse=True)
    first = torch.full((u.numel(),), x.numel(), dtype=torch.int64, device=x.device)
    first.scatter_reduce_(0, inv, torch.arange(x.numel(), device=x.device), "amin", include_self=True)
    return x[torch.sort(first).values]


def sample_neighbors(g: GNNGraph, nodes, K: int = -1, *, dir: str = "in", replace: bool = False,
                     dropnodes: bool = False, seed: Optional[int] = None) -> GNNGraph:
    """GNNGraphs/src/sampling.jl:68-121.  The result carries the edge feature `EID` (ids in `g`) and, with dropnodes,
    the node feature `NID`."""
    eids, _ = sample_edge_ids(g, nodes, K, dir=dir, replace=replace, seed=seed)
    dev = eids.device
    e0 = eids - 1
    s, t = g.s.to(dev)[e0], g.t.to(dev)[e0]
    w = None if g.w is None else g.w.to(dev)[e0]
    edata = {k: _take_last(x.to(dev), e0) for k, x in g.edata.items()}
    edata["EID"] = eids
    if not dropnodes:
        return GNNGraph(s, t, w, num_nodes=g.num_nodes, ndata=g.ndata, edata=edata, gdata=g.gdata,
                        num_graphs=g.num_graphs, graph_indicator=g.graph_indicator)
    nodes = _as_index(nodes).to(dev).reshape(-1).to(s.dtype)
    other = s if dir == "in" else t
    mark = torch.zeros(g.num_nodes + 1, dtype=torch.bool, device=dev)
    mark[nodes.long()] = True
    extra = _stable_unique(other[~mark[other.long()]])                  # setdiff(s, nodes): order of first appearance
    nodes_all = torch.cat([nodes, extra])
    nodemap = torch.zeros(g.num_nodes + 1, dtype=s.dtype, device=dev)
    nodemap[nodes_all.long()] = torch.arange(1, nodes_all.numel() + 1, dtype=s.dtype, device=dev)
    n0 = nodes_all.long() - 1
    ndata = {k: _take_last(x.to(dev), n0) for k, x in g.ndata.items()}
    ndata["NID"] = nodes_all
    gi = None if g.graph_indicator is None else g.graph_indicator.to(dev)[n0]
    return GNNGraph(nodemap[s.long()], nodemap[t.long()], w, num_nodes=int(nodes_all.numel()), ndata=ndata,
                    edata=edata, gdata=g.gdata, num_graphs=g.num_graphs, graph_indicator=gi)


def induced_subgraph(g: GNNGraph, nodes) -> GNNGraph:
    """GNNGraphs/src/sampling.jl:172-204: the nodes `nodes` (relabelled 1..len in the order given), every edge of `g`
    between two of them, node and edge features sliced.  Edges come grouped by target in the order of `nodes`, within
    a target in COO order — as the reference's loop over `neighbors(graph, node, dir = :in)` produces them."""
    nodes = _as_index(nodes).reshape(-1)
    if nodes.numel() == 0:
        return GNNGraph(torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64), num_nodes=0)
    eids, _ = sample_edge_ids(g, nodes, -1, dir="in")                   # all in-edges of the chosen targets
    dev = eids.device
    nodes = nodes.to(dev)
    e0 = eids - 1
    s, t = g.s.to(dev)[e0], g.t.to(dev)[e0]
    nodemap = torch.zeros(g.num_nodes + 1, dtype=s.dtype, device=dev)
    nodemap[nodes.long()] = torch.arange(1, nodes.numel() + 1, dtype=s.dtype, device=dev)
    keep = (nodemap[s.long()] > 0).nonzero().reshape(-1)
    e0 = e0[keep]
    n0 = nodes.long() - 1
    return GNNGraph(nodemap[s[keep].long()], nodemap[t[keep].long()], None if g.w is None else g.w.to(dev)[e0],
                    num_nodes=int(nodes.numel()), ndata={k: _take_last(x.to(dev), n0) for k, x in g.ndata.items()},
                    edata={k: _take_last(x.to(dev), e0) for k, x in g.edata.items()})


class NeighborLoader:
    """Mini-batches of sampled neighbourhoods (GNNGraphs/src/samplers.jl:28-105): for every batch of input nodes,
    `num_layers` rounds of "sample up to num_neighbors[layer] in-neighbours (with replacement, like the reference's
    `rand(neighbors, k)`) of the current frontier", then the subgraph induced by everything reached."""

    def __init__(self, graph: GNNGraph, *, num_neighbors: Sequence[int], input_nodes=None, num_layers: int,
                 batch_size: Optional[int] = None, seed: Optional[int] = None):
        self.graph = graph
        self.num_neighbors = list(num_neighbors)
        self.input_nodes = (torch.arange(1, graph.num_nodes + 1) if input_nodes is None
                            else _as_index(input_nodes).reshape(-1))
        self.num_layers = int(num_layers)
        self.batch_size = int(batch_size) if batch_size is not None else int(self.input_nodes.numel())
        self.seed = seed
        assert len(self.num_neighbors) >= self.num_layers

    def __len__(self):
        n = int(self.input_nodes.numel())
        return (n + self.batch_size - 1) // max(self.batch_size, 1) if n else 0

    def __iter__(self):
        g = self.graph
        n = int(self.input_nodes.numel())
        for b, start in enumerate(range(0, n, max(self.batch_size, 1))):
            batch = self.input_nodes[start:start + self.batch_size]
            dev = g.plan().device
            reached = batch.to(dev)
            frontier = reached
            for layer in range(self.num_layers):
                k = int(self.num_neighbors[layer])
                if k <= 0 or frontier.numel() == 0:
                    break
                sd = None if self.seed is None else self.seed + 1000003 * b + layer
                eids, offsets = sample_edge_ids(g, frontier, k, dir="in", replace=True, seed=sd)
                # the reference takes min(k, deg) draws per node: keep the first min(k, deg) of each node's k draws
                deg = _graph.degree(g, dir="in", edge_weight=False).to(dev)[frontier.long() - 1].to(torch.int64)
                pos = torch.arange(eids.numel(), device=dev) - offsets[:-1].repeat_interleave(offsets[1:] - offsets[:-1])
                owner = torch.repeat_interleave(torch.arange(frontier.numel(), device=dev), offsets[1:] - offsets[:-1])
                eids = eids[pos < torch.clamp(deg, max=k)[owner]]
                frontier = _stable_unique(g.s.to(dev)[eids - 1])
                reached = _stable_unique(torch.cat([reached, frontier]))
            yield induced_subgraph(g, reached)
