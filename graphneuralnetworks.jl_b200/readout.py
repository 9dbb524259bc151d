"""Graph-level readout on the same segment machinery:

    reduce_nodes / reduce_edges        GNNlib/src/utils.jl:12-42
    softmax_nodes / softmax_edges      GNNlib/src/utils.jl:44-72
    broadcast_nodes / broadcast_edges  GNNlib/src/utils.jl:105-121
    global_pool, global_attention_pool GNNlib/src/layers/pool.jl:3-12

`NNlib.scatter(aggr, x, graph_indicator)` is a segmented reduce whose "edges" are the nodes and whose "targets" are the
graphs: a bipartite plan (gnnb_graph_create with num_src = #items, num_dst = #graphs) lets the library's scatter /
gather / neighbourhood-softmax kernels (and their pullbacks) do all of it — no new kernels.
"""
from __future__ import annotations

import ctypes as C
import operator

import torch

from . import _lib
from . import graph as _graph
from ._lib import lib
from .graph import GNNGraph, _Plan, _stream, graph_indicator, rows, unrows
from .msgpass import _EdgeSoftmaxFn, _GatherFn, _ScatterFn, _aggr_code, _f32


class _IndicatorPlan:
    """plan of the bipartite graph  item k -> segment indicator[k]  (1-based, like graph_indicator)"""

    def __init__(self, indicator: torch.Tensor, num_segments: int, device):
        self.n_items = int(indicator.numel())
        self.n_segments = int(num_segments)
        ind = indicator.to(device=device, dtype=torch.int64).contiguous()
        src = torch.arange(1, self.n_items + 1, dtype=torch.int64, device=device)
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.gnnb_graph_create(C.byref(h), src.data_ptr(), ind.data_ptr(), self.n_items, self.n_items,
                                             self.n_segments, 8, 1, 1, _stream(device)))
        self.plan = _Plan(h.value, torch.device(device))


def _indicator_plan(g: GNNGraph, edges: bool) -> _IndicatorPlan:
    key = "_gi_plan_e" if edges else "_gi_plan_n"
    p = getattr(g, key, None)
    if p is None:
        dev = g.plan().device
        p = _IndicatorPlan(graph_indicator(g, edges=edges), g.num_graphs, dev)
        setattr(g, key, p)
    return p


def _reduce(aggr, ip: _IndicatorPlan, x: torch.Tensor) -> torch.Tensor:
    assert x.shape[-1] == ip.n_items
    r = _ScatterFn.apply(_f32(rows(x), ip.plan.device), ip.plan, _lib.DST, _aggr_code(aggr), ip.n_segments)
    return unrows(r)


def _broadcast(ip: _IndicatorPlan, x: torch.Tensor) -> torch.Tensor:
    assert x.shape[-1] == ip.n_segments
    r = _GatherFn.apply(_f32(rows(x), ip.plan.device), ip.plan, _lib.DST, ip.n_items)
    return unrows(r)


def reduce_nodes(aggr, g, x: torch.Tensor) -> torch.Tensor:
    """reduce_nodes(aggr, g, x) and reduce_nodes(aggr, indicator, x) — GNNlib/src/utils.jl:12-29."""
    if isinstance(g, GNNGraph):
        assert x.shape[-1] == g.num_nodes
        return _reduce(aggr, _indicator_plan(g, False), x)
    ind = g
    dev = _graph._compute_device(x)
    return _reduce(aggr, _IndicatorPlan(ind, int(ind.max()), dev), x)


def reduce_edges(aggr, g: GNNGraph, e: torch.Tensor) -> torch.Tensor:
    """GNNlib/src/utils.jl:31-42."""
    assert e.shape[-1] == g.num_edges
    return _reduce(aggr, _indicator_plan(g, True), e)


def softmax_nodes(g: GNNGraph, x: torch.Tensor) -> torch.Tensor:
    """Graph-wise softmax of the node features — GNNlib/src/utils.jl:44-57 (fused neighbourhood-softmax kernel)."""
    assert x.shape[-1] == g.num_nodes
    ip = _indicator_plan(g, False)
    return unrows(_EdgeSoftmaxFn.apply(_f32(rows(x), ip.plan.device), ip.plan))


def softmax_edges(g: GNNGraph, e: torch.Tensor) -> torch.Tensor:
    """Graph-wise softmax of the edge features — GNNlib/src/utils.jl:59-72: the reference's own sequence, including the
    `den .+ eps(eltype(e))` it adds only here."""
    assert e.shape[-1] == g.num_edges
    ip = _indicator_plan(g, True)
    mx = _broadcast(ip, _reduce(max, ip, e))
    num = torch.exp(e - mx)
    den = _broadcast(ip, _reduce(operator.add, ip, num))
    return num / (den + torch.finfo(e.dtype).eps)


def broadcast_nodes(g: GNNGraph, x: torch.Tensor) -> torch.Tensor:
    """GNNlib/src/utils.jl:105-110."""
    assert x.shape[-1] == g.num_graphs
    return _broadcast(_indicator_plan(g, False), x)


def broadcast_edges(g: GNNGraph, x: torch.Tensor) -> torch.Tensor:
    """GNNlib/src/utils.jl:116-121."""
    assert x.shape[-1] == g.num_graphs
    return _broadcast(_indicator_plan(g, True), x)


def global_pool(l, g: GNNGraph, x: torch.Tensor) -> torch.Tensor:
    """GNNlib/src/layers/pool.jl:3-5."""
    return reduce_nodes(l.aggr, g, x)


def global_attention_pool(l, g: GNNGraph, x: torch.Tensor) -> torch.Tensor:
    """GNNlib/src/layers/pool.jl:7-12."""
    alpha = softmax_nodes(g, l.fgate(x))
    feats = alpha * l.ffeat(x)
    return reduce_nodes(operator.add, g, feats)
