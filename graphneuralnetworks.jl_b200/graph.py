"""GNNGraph — the COO graph value of the reference (GNNGraphs/src/gnngraph.jl:108-117) plus the handful of
GNNGraphs functions that sit on the hot path:

    edge_index          GNNGraphs/src/query.jl:12-14
    degree              GNNGraphs/src/query.jl:314-369
    add_self_loops      GNNGraphs/src/transform.jl:12-28
    set_edge_weight     GNNGraphs/src/transform.jl:568-577
    batch (COO)         GNNGraphs/src/transform.jl:682-709
    graph_indicator     GNNGraphs/src/query.jl:500-512

Conventions follow the reference: node ids are 1-based Int64 (or Int32) vectors ``s`` (source) and ``t``
(target); feature arrays are Julia-shaped ``(D, num_nodes)`` / ``(K, num_edges)`` whose *memory* is
column-major (``colmajor`` below), i.e. every node owns D contiguous floats — exactly what a ``CuArray``
handed through ``ccall`` looks like to libgnnb200.

Each graph lazily owns one device *plan* (``gnnb_graph_t``: CSR by target, CSR by source on demand) that is
built once and cached — the reference rebuilds its CSC from COO on every fused call
(GNNGraphs/src/query.jl:227).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import lib


# --------------------------------------------------------------------------------------------------
# Julia-layout helpers
# --------------------------------------------------------------------------------------------------
def colmajor(x: torch.Tensor) -> torch.Tensor:
    """Return ``x`` (any strides) as a tensor of the same shape whose memory is column-major (Julia)."""
    nd = x.dim()
    if nd <= 1:
        return x.contiguous()
    rev = tuple(range(nd - 1, -1, -1))
    return x.permute(rev).contiguous().permute(rev)


def jl_zeros(*shape, dtype=torch.float32, device=None) -> torch.Tensor:
    rev = tuple(reversed(shape))
    return torch.zeros(rev, dtype=dtype, device=device).permute(tuple(range(len(shape) - 1, -1, -1)))


def jl_randn(*shape, dtype=torch.float32, device=None, generator=None) -> torch.Tensor:
    rev = tuple(reversed(shape))
    return torch.randn(rev, dtype=dtype, device=device, generator=generator).permute(
        tuple(range(len(shape) - 1, -1, -1)))


def rows(x: torch.Tensor) -> torch.Tensor:
    """(d1,...,dk, N) Julia array -> C-contiguous (N, dk,...,d1) view (copy only if x is not column-major)."""
    nd = x.dim()
    rev = tuple(range(nd - 1, -1, -1))
    r = x.permute(rev)
    return r if r.is_contiguous() else r.contiguous()


def unrows(r: torch.Tensor) -> torch.Tensor:
    """inverse of rows(): C-contiguous (N, dk,...,d1) -> Julia-shaped (d1,...,dk,N) column-major view."""
    nd = r.dim()
    return r.permute(tuple(range(nd - 1, -1, -1)))


def _compute_device(t: torch.Tensor) -> torch.device:
    """where the library computes for data that lives with `t`: t's GPU, else the current CUDA device (host arrays are
    staged there; there is no CPU path)."""
    if t.is_cuda:
        return t.device
    return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)


def _stream(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Plan:
    """Owner of one gnnb_graph_t."""

    def __init__(self, handle: int, device: torch.device):
        self.h = C.c_void_p(handle)
        self.device = device

    def __del__(self):
        try:
            if self.h:
                lib.gnnb_graph_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _as_index(v, device=None) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        t = v
    else:
        t = torch.as_tensor(np.asarray(v))
    if t.dtype not in (torch.int64, torch.int32):
        if t.is_floating_point() or t.dtype == torch.bool:
            raise ValueError("edge indices must be integers")
        t = t.to(torch.int64)
    if device is not None:
        t = t.to(device)
    return t.contiguous()


class GNNGraph:
    """COO graph ``(s, t[, w])`` with 1-based node ids — GNNGraph{<:COO_T} (GNNGraphs/src/gnngraph.jl:108-117).

    ``GNNGraph(s, t)``, ``GNNGraph((s, t))``, ``GNNGraph((s, t, w))`` and ``GNNGraph(adjacency_matrix)`` are
    accepted (gnngraph.jl:120-199); ``num_nodes`` defaults to ``max(maximum(s), maximum(t))``
    (convert.jl:33-36).  Indices are validated once, on the device, when the plan is built
    (``1 <= idx <= num_nodes``, convert.jl:49-54 -> AssertionError).
    """

    def __init__(self, s, t=None, w=None, *, num_nodes: Optional[int] = None, ndata=None, edata=None,
                 gdata=None, num_graphs: int = 1, graph_indicator=None, device=None):
        if t is None:
            as_edges = isinstance(s, tuple) and len(s) in (2, 3)           # (s, t) / (s, t, w): tuples
            if not as_edges and isinstance(s, list) and len(s) in (2, 3) and all(hasattr(v, "__len__") for v in s):
                # a nested list is an adjacency matrix when it is square ([[0,1],[1,0]]), a list of index vectors otherwise
                as_edges = not all(len(v) == len(s) for v in s)
            if as_edges:
                if len(s) == 3:
                    s, t, w = s
                else:
                    s, t = s
            else:  # adjacency matrix: A[i,j] != 0 <=> edge i -> j  (convert.jl:75-95, column-major findnz order)
                A = s if isinstance(s, torch.Tensor) else torch.as_tensor(np.asarray(s))
                if A.dim() != 2 or A.shape[0] != A.shape[1]:
                    raise ValueError("adjacency matrix must be square")
                nz = (A.t() != 0).nonzero()  # iterate columns first, like Julia's findnz on a dense matrix
                t, s = nz[:, 0] + 1, nz[:, 1] + 1
                if num_nodes is None:
                    num_nodes = A.shape[0]
                w = A.t()[A.t() != 0].to(torch.float32)        # v = A[nz] always travels as the edge weight (convert.jl:85)
        self.s = _as_index(s, device)
        self.t = _as_index(t, device)
        assert self.s.dim() == 1 and self.s.shape == self.t.shape, "s and t must be vectors of equal length"
        self.num_edges = int(self.s.numel())
        if num_nodes is None:
            num_nodes = int(max(int(self.s.max()), int(self.t.max()))) if self.num_edges else 0
        self.num_nodes = int(num_nodes)
        self.w = None if w is None else torch.as_tensor(w, dtype=torch.float32).to(self.s.device).contiguous()
        if self.w is not None:
            assert self.w.numel() == self.num_edges, "edge weight length must equal num_edges"  # convert.jl:47
        self.num_graphs = int(num_graphs)
        self.graph_indicator = graph_indicator
        self.ndata = dict(ndata or {}) if not isinstance(ndata, torch.Tensor) else {"x": ndata}
        self.edata = dict(edata or {}) if not isinstance(edata, torch.Tensor) else {"e": edata}
        self.gdata = dict(gdata or {}) if not isinstance(gdata, torch.Tensor) else {"u": gdata}
        for k, v in self.ndata.items():
            assert v.shape[-1] == self.num_nodes, f"ndata[{k}] last dim must be num_nodes"
        for k, v in self.edata.items():
            assert v.shape[-1] == self.num_edges, f"edata[{k}] last dim must be num_edges"
        self._plan: Optional[_Plan] = None
        self._loops: Optional["GNNGraph"] = None

    # -- conveniences mirroring g.x / g.e property access (datastore.jl getproperty)
    @property
    def x(self):
        return self.ndata["x"]

    @property
    def e(self):
        return self.edata["e"]

    @property
    def device(self) -> torch.device:
        return self.s.device

    def to(self, device) -> "GNNGraph":
        device = torch.device(device)
        g = GNNGraph(self.s.to(device), self.t.to(device), None if self.w is None else self.w.to(device),
                     num_nodes=self.num_nodes,
                     ndata={k: v.to(device) for k, v in self.ndata.items()},
                     edata={k: v.to(device) for k, v in self.edata.items()},
                     gdata={k: v.to(device) for k, v in self.gdata.items()},
                     num_graphs=self.num_graphs,
                     graph_indicator=None if self.graph_indicator is None else self.graph_indicator.to(device))
        return g

    def cuda(self) -> "GNNGraph":
        return self.to("cuda")

    def __repr__(self):
        return f"GNNGraph(num_nodes={self.num_nodes}, num_edges={self.num_edges}, num_graphs={self.num_graphs})"

    # -- the device plan ---------------------------------------------------------------------------
    def plan(self, device: Optional[torch.device] = None) -> _Plan:
        """Build (once) and return the device plan.  Raises AssertionError on out-of-range indices."""
        if self._plan is not None:
            return self._plan
        if device is None:
            device = _compute_device(self.s)
        if _lib.device_count() <= 0:
            raise _lib.GNNBError(_lib.ECUDA, "no CUDA device: the message-passing engine has no CPU fallback")
        h = C.c_void_p()
        on_dev = 1 if self.s.is_cuda else 0
        with torch.cuda.device(device):
            _lib.check(lib.gnnb_graph_create(C.byref(h), self.s.data_ptr(), self.t.data_ptr(), self.num_edges,
                                             self.num_nodes, self.num_nodes, self.s.element_size(), 1, on_dev,
                                             _stream(device)))
        self._plan = _Plan(h.value, device)
        return self._plan


# --------------------------------------------------------------------------------------------------
# queries / transforms on the hot path
# --------------------------------------------------------------------------------------------------
def edge_index(g: GNNGraph):
    """(s, t) — GNNGraphs/src/query.jl:12."""
    return g.s, g.t


def get_edge_weight(g: GNNGraph):
    return g.w


def set_edge_weight(g: GNNGraph, w: torch.Tensor) -> GNNGraph:
    """GNNGraphs/src/transform.jl:568-577."""
    assert w.numel() == g.num_edges
    h = GNNGraph(g.s, g.t, w, num_nodes=g.num_nodes, ndata=g.ndata, edata=g.edata, gdata=g.gdata,
                 num_graphs=g.num_graphs, graph_indicator=g.graph_indicator)
    h._plan = g._plan  # same topology: share the plan
    return h


def add_self_loops(g: GNNGraph) -> GNNGraph:
    """s=[s;1:n], t=[t;1:n], weights padded with 1 — GNNGraphs/src/transform.jl:12-28.

    Requires empty edata (the reference asserts it).  The result (and its plan, derived on the device from
    this graph's CSR without a new sort) is cached on ``g``: graphs are immutable values."""
    assert len(g.edata) == 0, "add_self_loops requires empty edata"  # transform.jl:14
    if g._loops is not None:
        return g._loops
    n = g.num_nodes
    nodes = torch.arange(1, n + 1, dtype=g.s.dtype, device=g.s.device)
    s = torch.cat([g.s, nodes])
    t = torch.cat([g.t, nodes])
    w = None if g.w is None else torch.cat([g.w, torch.ones(n, dtype=g.w.dtype, device=g.w.device)])
    h = GNNGraph(s, t, w, num_nodes=n, ndata=g.ndata, edata=g.edata, gdata=g.gdata, num_graphs=g.num_graphs,
                 graph_indicator=g.graph_indicator)
    if _lib.device_count() > 0:
        p = g.plan()
        hh = C.c_void_p()
        with torch.cuda.device(p.device):
            _lib.check(lib.gnnb_graph_add_self_loops(p.h, C.byref(hh), _stream(p.device)))
        h._plan = _Plan(hh.value, p.device)
    g._loops = h
    return h


def degree(g: GNNGraph, T=None, *, dir: str = "out", edge_weight=True) -> torch.Tensor:
    """degree(g, T; dir, edge_weight) — GNNGraphs/src/query.jl:314-331,355-369 (note the reference default dir=:out).

    edge_weight: True -> the graph's own weights if any, False/None -> counts, tensor -> those weights."""
    assert dir in ("in", "out", "both")  # query.jl:339
    if isinstance(edge_weight, torch.Tensor):
        w = edge_weight
    elif edge_weight is True:
        w = g.w
    else:
        w = None
    p = g.plan()
    out = torch.empty(g.num_nodes, dtype=torch.float32, device=p.device)
    if w is not None:
        assert w.numel() == g.num_edges
        w = w.to(device=p.device, dtype=torch.float32).contiguous()
    d = {"out": _lib.DIR_OUT, "in": _lib.DIR_IN, "both": _lib.DIR_BOTH}[dir]
    with torch.cuda.device(p.device):
        _lib.check(lib.gnnb_degree(p.h, d, _ptr(w), out.data_ptr(), _stream(p.device)))
    if T is None:
        T = torch.float32 if w is not None else g.s.dtype
    return out.to(T)


def graph_indicator(g: GNNGraph, edges: bool = False) -> torch.Tensor:
    """GNNGraphs/src/query.jl:500-512."""
    gi = g.graph_indicator
    if gi is None:
        gi = torch.ones(g.num_nodes, dtype=torch.int64, device=g.s.device)
    if edges:
        gi = gi[g.s.long() - 1]
    return gi


def batch(graphs: Sequence[GNNGraph]) -> GNNGraph:
    """Block-diagonal batching of COO graphs — GNNGraphs/src/transform.jl:682-709: node ids are offset by the
    cumulative node counts, graph_indicator by the cumulative graph counts; edges stay grouped per graph."""
    graphs = list(graphs)
    assert len(graphs) > 0
    dev = graphs[0].s.device
    nodesum = np.cumsum([0] + [g.num_nodes for g in graphs])
    graphsum = np.cumsum([0] + [g.num_graphs for g in graphs])
    s = torch.cat([g.s + int(nodesum[i]) for i, g in enumerate(graphs)])
    t = torch.cat([g.t + int(nodesum[i]) for i, g in enumerate(graphs)])
    ws = [g.w for g in graphs]
    w = None if any(x is None for x in ws) else torch.cat(ws)
    gi = torch.cat([graph_indicator(g) + int(graphsum[i]) for i, g in enumerate(graphs)]).to(dev)

    def cat(ds):
        keys = ds[0].keys()
        return {k: torch.cat([colmajor(d[k]) for d in ds], dim=-1) for k in keys}

    return GNNGraph(s, t, w, num_nodes=int(nodesum[-1]), ndata=cat([g.ndata for g in graphs]),
                    edata=cat([g.edata for g in graphs]), num_graphs=int(graphsum[-1]), graph_indicator=gi)


def rmat_graph(num_nodes: int, num_edges: int, seed: int = 17, device="cuda") -> GNNGraph:
    """Synthetic RMAT graph (ours — the reference has no RMAT generator): Graph500 parameters,
    counter-based splitmix64, generated on the device; bit-identical to oracle.orc_rmat."""
    device = torch.device(device)
    s = torch.empty(num_edges, dtype=torch.int64, device=device)
    t = torch.empty(num_edges, dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.gnnb_rmat_edges(num_nodes, num_edges, seed, s.data_ptr(), t.data_ptr(), _stream(device)))
    return GNNGraph(s, t, num_nodes=num_nodes)
