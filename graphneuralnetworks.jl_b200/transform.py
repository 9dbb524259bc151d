"""Edge-list transforms either side of the hot path, on the device:

    sort_edge_index(u, v)            GNNGraphs/src/utils.jl:41-45   (the reference's CUDA extension round-trips through
                                     the host: GNNGraphs/ext/GNNGraphsCUDAExt.jl:24-30)
    remove_self_loops(g)             GNNGraphs/src/transform.jl:49-64
    remove_multi_edges(g; aggr=+)    GNNGraphs/src/transform.jl:157-190
    to_bidirected(g)                 GNNGraphs/src/transform.jl:495-510
    unbatch(g)                       GNNGraphs/src/transform.jl:741-778
    csr(g; transposed)               the plan's COO -> CSR conversion as an API (the reference has no CSR type)

The index work (pair encoding, stable radix sort, duplicate runs) is csrc/transform.cu; the feature aggregation of
`remove_multi_edges` is the library's segmented scatter over the run ids it returns — the same kernels as
`aggregate_neighbors`.  Graphs given on the CPU are staged to the current CUDA device and the result lives there.
"""
from __future__ import annotations

import ctypes as C
import operator
from typing import List

import torch

from . import _lib
from . import graph as _graph
from . import readout as _readout
from ._lib import lib
from .graph import GNNGraph, _as_index, _stream, rows, unrows
from .msgpass import mean


def _on_device(g: GNNGraph) -> GNNGraph:
    dev = _graph._compute_device(g.s)
    return g if g.s.device == dev else g.to(dev)


def _take_edges(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """getobs(x, idx) on a Julia-shaped edge array (last dimension = edges)"""
    return unrows(rows(x)[idx])


def sort_edge_index(u, v=None, *, return_perm: bool = False):
    """Lexicographic, stable sort of the pairs (u[k], v[k]) — `sort_edge_index(u, v)` / `sort_edge_index((u, v))`.
    Returns (u_sorted, v_sorted) and, with return_perm, the 0-based permutation as a third value."""
    if v is None:
        u, v = u
    u, v = _as_index(u), _as_index(v)
    assert u.dim() == 1 and u.shape == v.shape, "u and v must be vectors of equal length"
    dev = _graph._compute_device(u)
    u, v = u.to(dev), v.to(device=dev, dtype=u.dtype)
    n = int(u.numel())
    uo, vo = torch.empty_like(u), torch.empty_like(v)
    perm = torch.empty(n, dtype=torch.int64, device=dev)
    if n:
        hi = int(max(int(u.max()), int(v.max())))
        with torch.cuda.device(dev):
            _lib.check(lib.gnnb_sort_edge_index(u.data_ptr(), v.data_ptr(), n, max(hi, 0), u.element_size(),
                                                uo.data_ptr(), vo.data_ptr(), perm.data_ptr(), _stream(dev)))
    return (uo, vo, perm) if return_perm else (uo, vo)


def remove_self_loops(g: GNNGraph) -> GNNGraph:
    """Drop the edges with s == t; edge weights and features follow (transform.jl:49-64)."""
    keep = (g.s != g.t).nonzero().reshape(-1)
    return GNNGraph(g.s[keep], g.t[keep], None if g.w is None else g.w[keep], num_nodes=g.num_nodes, ndata=g.ndata,
                    edata={k: _take_edges(x, keep.to(x.device)) for k, x in g.edata.items()}, gdata=g.gdata,
                    num_graphs=g.num_graphs, graph_indicator=g.graph_indicator)


def remove_multi_edges(g: GNNGraph, aggr=operator.add) -> GNNGraph:
    """One edge per distinct (s, t), in (s, t) order; weights and edge features of the collapsed edges are reduced with
    `aggr` (+, mean, max, min) — transform.jl:157-190."""
    g = _on_device(g)
    dev, E = g.s.device, g.num_edges
    if E == 0:
        return g
    so, to = torch.empty_like(g.s), torch.empty_like(g.t)
    perm = torch.empty(E, dtype=torch.int64, device=dev)
    seg = torch.empty(E, dtype=torch.int64, device=dev)
    nu = C.c_int64(0)
    with torch.cuda.device(dev):
        _lib.check(lib.gnnb_coalesce_edges(g.s.data_ptr(), g.t.data_ptr(), E, g.num_nodes, g.s.element_size(), 1,
                                           so.data_ptr(), to.data_ptr(), perm.data_ptr(), seg.data_ptr(), C.byref(nu),
                                           _stream(dev)))
    nu = int(nu.value)
    ip = _readout._IndicatorPlan(seg, nu, dev)        # sorted edge k -> distinct pair seg[k]: `_scatter(aggr, ·, idxs)`

    def reduce(x: torch.Tensor) -> torch.Tensor:
        return _readout._reduce(aggr, ip, _take_edges(x.to(dev), perm))

    return GNNGraph(so[:nu].clone(), to[:nu].clone(), None if g.w is None else reduce(g.w), num_nodes=g.num_nodes,
                    ndata=g.ndata, edata={k: reduce(x) for k, x in g.edata.items()}, gdata=g.gdata,
                    num_graphs=g.num_graphs, graph_indicator=g.graph_indicator)


def to_bidirected(g: GNNGraph) -> GNNGraph:
    """Add the reverse of every edge, then remove_multi_edges with mean (transform.jl:495-510)."""
    both = GNNGraph(torch.cat([g.s, g.t]), torch.cat([g.t, g.s]), None if g.w is None else torch.cat([g.w, g.w]),
                    num_nodes=g.num_nodes, ndata=g.ndata,
                    edata={k: unrows(torch.cat([rows(x), rows(x)], dim=0)) for k, x in g.edata.items()},
                    gdata=g.gdata, num_graphs=g.num_graphs, graph_indicator=g.graph_indicator)
    return remove_multi_edges(both, aggr=mean)


def unbatch(g: GNNGraph) -> List[GNNGraph]:
    """Split a batched graph back into its components (transform.jl:741-778).  Node ids are shifted back, node/edge
    features are sliced; the edges must be grouped per graph (as `batch` leaves them) — asserted like the reference."""
    if g.num_graphs == 1:
        return [g]
    gi = g.graph_indicator.to(torch.int64).cpu()
    assert bool((gi[1:] >= gi[:-1]).all()), "The graph_indicator vector must be sorted."
    n_per = torch.bincount(gi - 1, minlength=g.num_graphs)
    cum = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(n_per, 0)])
    ge = gi[(g.s.to(torch.int64) - 1).cpu()] - 1                       # graph of each edge (by its source)
    assert bool((ge[1:] >= ge[:-1]).all()), \
        "Error in unbatching, likely the edges are not sorted (first edges belong to the first graphs, then edges in " \
        "the second graph and so on)"
    e_per = torch.bincount(ge, minlength=g.num_graphs)
    ecum = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(e_per, 0)])
    out = []
    for i in range(g.num_graphs):
        n0, n1, e0, e1 = int(cum[i]), int(cum[i + 1]), int(ecum[i]), int(ecum[i + 1])
        out.append(GNNGraph(g.s[e0:e1] - n0, g.t[e0:e1] - n0, None if g.w is None else g.w[e0:e1], num_nodes=n1 - n0,
                            ndata={k: x[..., n0:n1] for k, x in g.ndata.items()},
                            edata={k: x[..., e0:e1] for k, x in g.edata.items()},
                            gdata={k: x[..., i] for k, x in g.gdata.items()}))
    return out


def csr(g: GNNGraph, transposed: bool = False):
    """(rowptr, col, eid) of the plan as int32 device tensors: CSR by target (col = source of each sorted edge), or by
    source with transposed=True; eid[k] = 0-based COO position of sorted edge k (stable within a row)."""
    p = g.plan()
    nrows = g.num_nodes
    rowptr = torch.empty(nrows + 1, dtype=torch.int32, device=p.device)
    col = torch.empty(g.num_edges, dtype=torch.int32, device=p.device)
    eid = torch.empty(g.num_edges, dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(lib.gnnb_graph_csr_device(p.h, int(bool(transposed)), rowptr.data_ptr(), col.data_ptr(),
                                             eid.data_ptr(), _stream(p.device)))
    return rowptr, col, eid
