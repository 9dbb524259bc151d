"""Node-partitioned message passing across the GPUs of one box (the reference has no
distributed code at all — this is new work prescribed by BASELINE.json's north_star).

Layout: nodes are split into `world` contiguous ranges (cost-balanced: a node costs NODE_COST edge-equivalents
for the dense per-node work plus its in- and out-degree).  Rank p owns x[:, lo_p:hi_p], every in-edge of its
nodes (forward shard) and every out-edge of its nodes (backward shard).  One pass =

    pack the rows each peer asked for (gnnb_gather_rows)  ->  one all-to-all-v over NCCL/NVLink
    ->  ONE fused segmented-reduce kernel over the [local rows | halo rows] source space (gnnb_propagate_halo)

The halo lists are deduplicated per peer (each remote row crosses NVLink once per pass) and built once.
Both directions are "pull": the backward pass gathers dout rows of remote targets through the backward shard,
so there is no scatter-reduce across GPUs and every result stays deterministic; inside a target row the
edges keep their COO order, so a shard reproduces the single-GPU summation order.

`torch.distributed` is plumbing (process group, all_to_all_single); index construction below is plain torch
ops that also run on CPU tensors with the gloo backend (tests/test_partition_gloo.py).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import lib
from .graph import _Plan, _ptr, _stream, rows, unrows

NODE_COST = 12  # dense per-node work (GEMM, bias, relu, grads) in edge-equivalents, from the 1-GPU profile (1.9 ns/node vs 0.17 ns/edge)


# ---------------------------------------------------------------------------------------------------------
# index construction (device agnostic)
# ---------------------------------------------------------------------------------------------------------
def balanced_bounds(cost: torch.Tensor, world: int) -> List[int]:
    """contiguous ranges [b[p], b[p+1]) with ~equal total cost."""
    n = cost.numel()
    cs = torch.cumsum(cost.to(torch.float64), 0)
    total = float(cs[-1]) if n else 0.0
    targets = torch.tensor([total * p / world for p in range(1, world)], dtype=torch.float64, device=cost.device)
    cuts = torch.searchsorted(cs, targets).tolist() if world > 1 else []
    b = [0] + [min(max(int(c) + 1, 0), n) for c in cuts] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def ownership_first(num_nodes: int, world: int, ownership: str, bounds: Optional[List[int]] = None) -> List[int]:
    """Start of every rank's range in partition-id space (world + 1 entries).  'contiguous': the node ranges themselves
    (`bounds`, or equal ranges); 'cyclic': node v (0-based) belongs to rank v % world as local row v // world, so rank q
    owns ceil((N - q) / world) nodes — the hubs of a skewed id space (RMAT: the low ids) spread over all ranks."""
    if ownership == "contiguous":
        return list(bounds) if bounds is not None else [(num_nodes * q) // world for q in range(world + 1)]
    assert ownership in ("cyclic", "balanced"), ownership
    first = [0]
    for q in range(world):
        first.append(first[-1] + (num_nodes - q + world - 1) // world)
    return first


def to_pid(v0: torch.Tensor, world: int, first: List[int], ownership: str, relabel: Optional[torch.Tensor] = None) -> torch.Tensor:
    """node id (0-based) -> partition id: the bijection onto [0, N) in which every rank owns one contiguous range"""
    if ownership == "contiguous":
        return v0
    if relabel is not None:
        v0 = relabel.to(v0.dtype)[v0]
    ft = torch.tensor(first[:-1], dtype=v0.dtype, device=v0.device)
    return ft[v0 % world] + v0 // world


def degree_order(chunks, num_nodes: int, device) -> torch.Tensor:
    """nodes by decreasing in+out degree (stable: ties keep id order) — the order in which 'balanced' ownership deals
    them to the ranks; the torch restatement of gnnb_degree_accumulate + the sort of gnnb_balanced_relabel."""
    cost = torch.zeros(num_nodes, dtype=torch.int64, device=device)
    for sc, tc in chunks:
        cost += torch.bincount(sc.to(device).to(torch.int64) - 1, minlength=num_nodes)
        cost += torch.bincount(tc.to(device).to(torch.int64) - 1, minlength=num_nodes)
    return torch.sort(cost, descending=True, stable=True).indices


def build_shard(key0: torch.Tensor, other0: torch.Tensor, lo: int, hi: int, bounds: List[int],
                self_loops: bool):
    """Edges whose reduction row `key0` (0-based partition id) lies in [lo,hi), re-indexed for one GPU — the torch
    restatement of csrc/shard.cu, used for CPU tensors (gloo tests of the host logic).

    Returns dict(row: local reduction row, col: gathered node in [local | halo] space, halo: sorted partition ids of
    the remote gathered nodes, halo_local: their owner-local rows, recv_counts: rows expected from every owner)."""
    sel = (key0 >= lo) & (key0 < hi)
    k = key0[sel] - lo
    o = other0[sel]
    n_local = hi - lo
    is_local = (o >= lo) & (o < hi)
    halo = torch.unique(o[~is_local])                       # sorted => grouped by owner (ranges are contiguous)
    col = torch.where(is_local, o - lo, n_local + torch.searchsorted(halo, o))
    if self_loops:                                          # (i,i) appended after the originals (transform.jl:17-19)
        loops = torch.arange(n_local, dtype=k.dtype, device=k.device)
        k = torch.cat([k, loops])
        col = torch.cat([col, loops])
    edges = torch.tensor(bounds[1:], dtype=halo.dtype, device=halo.device)
    owner = torch.bucketize(halo, edges, right=True)
    recv_counts = torch.bincount(owner, minlength=len(bounds) - 1).tolist()
    starts = torch.tensor(bounds[:-1], dtype=halo.dtype, device=halo.device)
    halo_local = (halo - starts[owner]).to(torch.int32)
    return {"row": k, "col": col, "halo": halo, "halo_local": halo_local, "recv_counts": recv_counts, "n_local": n_local}


def exchange_requests(halo_local: torch.Tensor, recv_counts: List[int], group=None):
    """Tell every owner which of its rows (owner-local int32 indices, grouped by owner) this rank needs.  Returns
    (send_idx: int32 local row ids to pack, in peer order; send_counts)."""
    dev = halo_local.device
    rc = torch.tensor(recv_counts, dtype=torch.int64, device=dev)
    sc = torch.empty_like(rc)
    dist.all_to_all_single(sc, rc, group=group)
    send_counts = sc.tolist()
    wanted = torch.empty(int(sum(send_counts)), dtype=torch.int32, device=dev)
    dist.all_to_all_single(wanted, halo_local.to(torch.int32).contiguous(), output_split_sizes=send_counts,
                           input_split_sizes=recv_counts, group=group)
    return wanted, send_counts


# ---------------------------------------------------------------------------------------------------------
# the distributed graph
# ---------------------------------------------------------------------------------------------------------
class _Shard:
    def __init__(self, n_local, n_halo, recv_counts, num_edges, send_idx, send_counts, plan):
        self.n_local = int(n_local)
        self.n_halo = int(n_halo)
        self.recv_counts = [int(v) for v in recv_counts]
        self.send_idx = send_idx
        self.send_counts = send_counts
        self.plan = plan
        self.num_edges = int(num_edges)
        self.push = None           # per-D state of the peer-to-peer push path (_PushState)


def rmat_chunks(num_nodes: int, num_edges: int, seed: int, device, chunk_edges: int = 1 << 26):
    """the counter-based RMAT list of gnnb_rmat_edges as (src, dst) int64 chunks (1-based ids) generated on `device`; the two buffers are
    reused, so a chunk is valid until the next one is requested"""
    dev = torch.device(device)
    cap = min(chunk_edges, max(num_edges, 1))
    s, t = (torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(2))
    for first in range(0, num_edges, chunk_edges):
        cnt = min(chunk_edges, num_edges - first)
        with torch.cuda.device(dev):
            _lib.check(lib.gnnb_rmat_edges_range(num_nodes, first, cnt, seed, s.data_ptr(), t.data_ptr(), _stream(dev)))
        yield s[:cnt], t[:cnt]


class DistGraph:
    """A GNNGraph partitioned over the ranks of `group`.

    DistGraph(s, t, num_nodes, ...): every rank passes the same global COO (1-based s, t).
    DistGraph.from_chunks(chunks, num_nodes, ...): the same from an iterable of (s, t) chunks — the global list is never
    resident (CUDA only).  DistGraph.from_rmat(...) generates the chunks of the counter-based RMAT list on the device.
    On a CUDA device the shards are built by csrc/shard.cu (stable scan-compaction of every chunk, sort + unique of the
    remote ids, renaming, plan); on CPU tensors (gloo tests) by the torch restatement above.

    ownership = 'contiguous' (node ranges, cost-balanced unless `bounds` is given), 'cyclic' (0-based node v on rank
    v % world) or 'balanced' (nodes sorted by decreasing degree and dealt to the ranks in turn: edges, nodes and served
    halo rows all balanced whatever the id space looks like — RMAT probabilities are products over id bits, so neither
    ranges nor v % world balance it; needs `chunks` to be iterable twice).  `local_nodes()` lists the owned node ids in
    local-row order."""

    def __init__(self, s: Optional[torch.Tensor], t: Optional[torch.Tensor], num_nodes: int, *, add_self_loops: bool = False,
                 group=None, device=None, bounds: Optional[List[int]] = None, ownership: str = "contiguous",
                 chunks=None, chunk_edges: int = 1 << 26):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.timing = {}                                     # build phases, ms (this rank)
        self.device = torch.device(device) if device is not None else s.device
        self.num_nodes = int(num_nodes)
        self.ownership = ownership
        self.self_loops = add_self_loops
        self._c = None
        if chunks is None:
            s = s.to(self.device)
            t = t.to(self.device)
            if ownership == "contiguous" and bounds is None:
                s0, t0 = s.to(torch.int64) - 1, t.to(torch.int64) - 1
                cost = (torch.bincount(t0, minlength=num_nodes) + torch.bincount(s0, minlength=num_nodes) + NODE_COST)
                bounds = balanced_bounds(cost, self.world)
                del s0, t0, cost
            E = int(s.numel())
            chunks = [(s[i:i + chunk_edges], t[i:i + chunk_edges]) for i in range(0, max(E, 1), chunk_edges)] if E else []
        self._order = self._relabel = None
        if ownership == "balanced":
            if not isinstance(chunks, (list, tuple)) and not callable(chunks):
                raise ValueError("ownership='balanced' needs the chunks twice: pass a list or a callable that returns an iterator")
            W, N = self.world, self.num_nodes
            it = chunks() if callable(chunks) else chunks
            t_own = time.perf_counter()
            if self.device.type == "cuda":                   # degree histogram + stable sort + deal, on the device
                cost = torch.zeros(N, dtype=torch.int32, device=self.device)
                self._relabel = torch.empty(N, dtype=torch.int32, device=self.device)
                order = torch.empty(N, dtype=torch.int32, device=self.device)
                with torch.cuda.device(self.device):
                    for sc, tc in it:
                        sc, tc = sc.to(self.device).contiguous(), tc.to(self.device).contiguous()
                        _lib.check(lib.gnnb_degree_accumulate(sc.data_ptr(), tc.data_ptr(), sc.numel(), sc.element_size(), 1, N,
                                                              cost.data_ptr(), _stream(self.device)))
                    _lib.check(lib.gnnb_balanced_relabel(cost.data_ptr(), N, W, self._relabel.data_ptr(), order.data_ptr(),
                                                         _stream(self.device)))
                self._order = order
                del cost
            else:                                            # the same deal in torch ops (gloo tests)
                by_degree = degree_order(it, N, self.device)
                pos = torch.arange(N, device=self.device)
                r, j = pos // W, pos % W
                o = torch.where((r % 2 == 1) & (r < N // W), W - 1 - j, j)
                self._relabel = torch.empty(N, dtype=torch.int32, device=self.device)
                self._relabel[by_degree] = (r * W + o).to(torch.int32)
                self._order = torch.empty(N, dtype=torch.int64, device=self.device)      # position -> node
                self._order[self._relabel.long()] = pos
                del by_degree, pos, r, j, o
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            self.timing["ownership_ms"] = (time.perf_counter() - t_own) * 1e3
        if callable(chunks):
            chunks = chunks()
        self.first = ownership_first(self.num_nodes, self.world, ownership, bounds)
        self.bounds = self.first
        self.lo, self.hi = self.first[self.rank], self.first[self.rank + 1]
        self.n_local = self.hi - self.lo
        if self.device.type == "cuda":
            self.fwd, self.bwd = self._build_native(chunks)
            torch.cuda.synchronize(self.device)
        else:
            self.fwd, self.bwd = self._build_torch(chunks)

    @classmethod
    def from_chunks(cls, chunks, num_nodes: int, **kw):
        return cls(None, None, num_nodes, chunks=chunks, **kw)

    @classmethod
    def from_rmat(cls, num_nodes: int, num_edges: int, seed: int = 17, *, device, chunk_edges: int = 1 << 26, **kw):
        """the RMAT list of gnnb_rmat_edges, generated (identically on every rank) and consumed chunk by chunk"""
        dev = torch.device(device)
        return cls(None, None, num_nodes, chunks=lambda: rmat_chunks(num_nodes, num_edges, seed, dev, chunk_edges), device=dev, **kw)

    def local_nodes(self) -> torch.Tensor:
        """0-based global node id of every local row, in local-row order"""
        if self.ownership == "contiguous":
            return torch.arange(self.lo, self.hi, device=self.device)
        pos = self.rank + self.world * torch.arange(self.n_local, device=self.device)
        return pos if self._order is None else self._order[pos].long()

    # -- shard construction on the device (csrc/shard.cu)
    def _build_native(self, chunks):
        dev, world = self.device, self.world
        b = C.c_void_p()
        bounds_arr = (C.c_int64 * (world + 1))(*self.first) if self.ownership == "contiguous" else None
        shards = []
        t_sh = time.perf_counter()
        with torch.cuda.device(dev):
            st = _stream(dev)
            _lib.check(lib.gnnb_shard_builder_create(C.byref(b), self.num_nodes, world, self.rank,
                                                     0 if self.ownership == "contiguous" else 1, bounds_arr,
                                                     None if self._relabel is None else self._relabel.data_ptr()))
            try:
                for sc, tc in chunks:
                    sc, tc = sc.to(dev).contiguous(), tc.to(dev).contiguous()
                    assert sc.dtype == tc.dtype and sc.dtype in (torch.int32, torch.int64)
                    _lib.check(lib.gnnb_shard_builder_add(b, sc.data_ptr(), tc.data_ptr(), sc.numel(), sc.element_size(), 1, st))
                for direction in (0, 1):
                    h = C.c_void_p()
                    nl, nh, ne = C.c_int64(), C.c_int64(), C.c_int64()
                    rc = (C.c_int64 * world)()
                    _lib.check(lib.gnnb_shard_builder_finish(b, direction, int(self.self_loops), C.byref(h), C.byref(nl),
                                                             C.byref(nh), C.byref(ne), rc, st))
                    halo_local = torch.empty(max(nh.value, 1), dtype=torch.int32, device=dev)[:nh.value]
                    _lib.check(lib.gnnb_shard_builder_halo(b, direction, halo_local.data_ptr() if nh.value else None, st))
                    shards.append((h, nl.value, nh.value, ne.value, list(rc), halo_local))
            finally:
                lib.gnnb_shard_builder_destroy(b)
        torch.cuda.synchronize(dev)
        self.timing["shards_ms"] = (time.perf_counter() - t_sh) * 1e3
        t_ex = time.perf_counter()
        out = []
        for h, nl, nh, ne, rc, halo_local in shards:
            send_idx, send_counts = exchange_requests(halo_local, rc, self.group)
            out.append(_Shard(nl, nh, rc, ne, send_idx, send_counts, _Plan(h.value, dev)))
        torch.cuda.synchronize(dev)
        self.timing["request_exchange_ms"] = (time.perf_counter() - t_ex) * 1e3
        return out

    def _plans(self, d):
        """plan of a torch-built shard dict over [local | halo] sources (the gloo tests call this with the CPU test double
        installed)"""
        h = C.c_void_p()
        col, row = d["col"].to(torch.int32).contiguous(), d["row"].to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(lib.gnnb_graph_create(C.byref(h), col.data_ptr(), row.data_ptr(), row.numel(),
                                             d["n_local"] + int(d["halo"].numel()), d["n_local"], 4, 0, 1, _stream(self.device)))
        return _Plan(h.value, self.device)

    # -- the same in torch ops, for CPU tensors (host logic under gloo)
    def _build_torch(self, chunks):
        chunks = list(chunks)
        s0 = torch.cat([c[0] for c in chunks]).to(self.device).to(torch.int64) - 1 if chunks else torch.zeros(0, dtype=torch.int64)
        t0 = torch.cat([c[1] for c in chunks]).to(self.device).to(torch.int64) - 1 if chunks else torch.zeros(0, dtype=torch.int64)
        ps = to_pid(s0, self.world, self.first, self.ownership, self._relabel)
        pt = to_pid(t0, self.world, self.first, self.ownership, self._relabel)
        out = []
        for key0, other0 in ((pt, ps), (ps, pt)):
            d = build_shard(key0, other0, self.lo, self.hi, self.first, self.self_loops)
            send_idx, send_counts = exchange_requests(d["halo_local"], d["recv_counts"], self.group)
            out.append(_Shard(d["n_local"], d["halo"].numel(), d["recv_counts"], d["row"].numel(), send_idx, send_counts,
                              None))
        return out

    # -- halo exchange: rows (n_local, D) -> halo rows (n_halo, D)
    def halo(self, shard: _Shard, x_rows: torch.Tensor) -> torch.Tensor:
        D = x_rows.shape[1]
        n_send = int(shard.send_idx.numel())
        send = torch.empty((n_send, D), dtype=x_rows.dtype, device=x_rows.device)
        if x_rows.is_cuda:
            with torch.cuda.device(self.device):
                _lib.check(lib.gnnb_gather_rows(shard.send_idx.data_ptr(), n_send, x_rows.data_ptr(), D,
                                                send.data_ptr(), _stream(self.device)))
        else:  # gloo/CPU test path for the host logic only
            send = x_rows[shard.send_idx.long()]
        recv = torch.empty((shard.n_halo, D), dtype=x_rows.dtype, device=x_rows.device)
        dist.all_to_all_single(recv, send, output_split_sizes=shard.recv_counts,
                               input_split_sizes=shard.send_counts, group=self.group)
        return recv

    # -- halo exchange, push path: one kernel writes the requested rows into every peer's halo buffer over NVLink
    def _push_state(self, shard: _Shard, D: int):
        """Peer-mapped double-buffered halo buffers for rows of D floats (built once per shard and D).  Every rank runs
        the same collectives whatever happens locally; if any rank fails to allocate / export / map, ALL ranks fall
        back to the NCCL exchange (returns None)."""
        if shard.push is None:
            shard.push = {}
        if D in shard.push:
            return shard.push[D]
        dev, world, rank = self.device, self.world, self.rank
        ok = 1
        all_recv = [None] * world
        dist.all_gather_object(all_recv, list(shard.recv_counts), group=self.group)
        row0 = [int(sum(all_recv[q][:rank])) for q in range(world)]   # where my rows start in every peer's halo buffer
        # double buffer: pass k+1 never overwrites what a slower rank's pass k still reads.  GNNB_HALO_BUFFERS=1 is safe when
        # passes over a shard alternate with passes over the other one (one layer, forward / backward: the other pass's
        # completion collective orders them) and halves the memory — what config 5's 1 KB rows need
        nbuf = 1 if os.environ.get("GNNB_HALO_BUFFERS", "2") == "1" else 2
        bufs, handles = [], []
        try:
            with torch.cuda.device(dev):
                for _ in range(nbuf):
                    ptr = C.c_void_p()
                    _lib.check(lib.gnnb_dev_alloc(C.byref(ptr), max(shard.n_halo, 1) * D * 4))
                    h = (C.c_ubyte * 64)()
                    _lib.check(lib.gnnb_ipc_get_handle(ptr, h))
                    bufs.append(ptr.value)
                    handles.append(bytes(h))
        except Exception:
            ok = 0
            handles = [bytes(64)] * nbuf
        all_handles = [None] * world
        dist.all_gather_object(all_handles, handles, group=self.group)
        peer_ptrs = [[0] * world for _ in range(nbuf)]
        try:
            with torch.cuda.device(dev):
                for q in range(world):
                    if q == rank or shard.send_counts[q] == 0:
                        continue
                    for b in range(nbuf):
                        pp = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(all_handles[q][b])
                        _lib.check(lib.gnnb_ipc_open_handle(hb, C.byref(pp)))
                        peer_ptrs[b][q] = pp.value
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:                               # NCCL for everyone: give back what this rank did set up
            self._release_push({"bufs": bufs, "peer_ptrs": peer_ptrs})
            shard.push[D] = None
            return None
        seg = [0]
        for q in range(world):
            seg.append(seg[-1] + int(shard.send_counts[q]))
        st = {"bufs": bufs, "row0": (C.c_int64 * world)(*row0), "seg": (C.c_int64 * (world + 1))(*seg),
              "peer_c": [(C.c_void_p * world)(*[C.c_void_p(v) for v in peer_ptrs[b]]) for b in range(nbuf)], "turn": 0,
              "nbuf": nbuf,
              "peer_ptrs": peer_ptrs,
              "flag": torch.zeros(1, device=dev)}
        shard.push[D] = st
        return st

    def _release_push(self, st) -> None:
        with torch.cuda.device(self.device):
            for per_buf in st["peer_ptrs"]:
                for pp in per_buf:
                    if pp:
                        lib.gnnb_ipc_close_handle(C.c_void_p(pp))
            for ptr in st["bufs"]:
                if ptr:
                    lib.gnnb_dev_free(C.c_void_p(ptr))

    def close(self) -> None:
        """unmap the peers' halo buffers and free this rank's (call on every rank once no pass is in flight)"""
        for sh in (getattr(self, "fwd", None), getattr(self, "bwd", None)):
            if sh is None or not sh.push:
                continue
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            for st in sh.push.values():
                if st is not None:
                    self._release_push(st)
            sh.push = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def halo_ptr(self, shard: _Shard, x_rows: torch.Tensor) -> int:
        """device pointer of this rank's halo rows for `x_rows` (valid until the next-but-one call for this shard)"""
        if self.world == 1 or os.environ.get("GNNB_HALO", "push") != "push":
            t = self.halo(shard, x_rows)
            self._keep = t                                      # keep the NCCL receive buffer alive for the kernel
            return t.data_ptr()
        D = x_rows.shape[1]
        st = self._push_state(shard, D)
        if st is None:                                          # some rank could not set up peer mapping: NCCL for everyone
            t = self.halo(shard, x_rows)
            self._keep = t
            return t.data_ptr()
        b = st["turn"]
        st["turn"] = (b + 1) % st["nbuf"]
        with torch.cuda.device(self.device):
            _lib.check(lib.gnnb_halo_push(shard.send_idx.data_ptr(), st["seg"], st["peer_c"][b], st["row0"], self.world,
                                          x_rows.data_ptr(), D, _stream(self.device)))
        dist.all_reduce(st["flag"], group=self.group)           # every rank's push kernel precedes its part of this collective
        return st["bufs"][b]

    def gcn_c(self):
        """c = 1/sqrt(in-degree) of the owned nodes (exact: rowptr differences of the forward shard), plus the
        halo copies the two shards need.  Computed once."""
        if self._c is None:
            p = self.fwd.plan
            c = torch.empty(self.n_local, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(lib.gnnb_gcn_norm(p.h, None, c.data_ptr(), _stream(self.device)))
            cf = torch.cat([c, self.halo(self.fwd, c.reshape(-1, 1)).reshape(-1)])
            cb = torch.cat([c, self.halo(self.bwd, c.reshape(-1, 1)).reshape(-1)])
            self._c = (c, cf.contiguous(), cb.contiguous())
        return self._c

    def propagate(self, shard: _Shard, x_rows: torch.Tensor, cs, ct, aggr=_lib.SUM) -> torch.Tensor:
        slices = int(os.environ.get("GNNB_HALO_SLICES", "1"))
        if slices > 1 and x_rows.shape[1] % slices == 0 and not getattr(self, "_slicing", False):
            # column-sliced pass: exchange and reduce `D / slices` feature columns at a time, so that the halo buffers
            # shrink by `slices` (config 5: 62 GB of halo rows per pass at D = 256 do not fit beside the features);
            # costs two strided copies of the local rows and re-reads the index arrays once per slice
            w = x_rows.shape[1] // slices
            out = torch.empty_like(x_rows)
            self._slicing = True
            try:
                for i in range(slices):
                    out[:, i * w:(i + 1) * w] = self.propagate(shard, x_rows[:, i * w:(i + 1) * w].contiguous(), cs, ct, aggr)
            finally:
                self._slicing = False
            return out
        D = x_rows.shape[1]
        hptr = self.halo_ptr(shard, x_rows)
        out = torch.empty_like(x_rows)
        with torch.cuda.device(self.device):
            _lib.check(lib.gnnb_propagate_halo(shard.plan.h, _lib.COPY_XJ, aggr, x_rows.data_ptr(),
                                               hptr if shard.n_halo else None, shard.n_local, None,
                                               _ptr(cs), _ptr(ct), D, out.data_ptr(), _stream(self.device)))
        return out


class _DistGCNPropagateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_rows, dg: DistGraph):
        c, cf, cb = dg.gcn_c()
        ctx.dg = dg
        return dg.propagate(dg.fwd, x_rows.contiguous(), cf, c)

    @staticmethod
    def backward(ctx, dout):
        dg = ctx.dg
        c, cf, cb = dg.gcn_c()
        return dg.propagate(dg.bwd, dout.contiguous(), cb, c), None


def dist_gcn_conv(l, dg: DistGraph, x_local: torch.Tensor) -> torch.Tensor:
    """gcn_conv (GNNlib/src/layers/conv.jl:14-72) on the rows this rank owns; x_local is Julia-shaped (Din, n_local).
    The graph must have been partitioned with add_self_loops = l.add_self_loops.  Weight gradients are per-rank
    partial sums: all-reduce them like any data-parallel layer."""
    assert dg.self_loops == bool(l.add_self_loops)
    from .layers import _linear
    W = l.weight
    Dout, Din = W.shape
    x = x_local
    if Dout < Din:
        x = _linear(l, W, x, False)
    pr = unrows(_DistGCNPropagateFn.apply(rows(x), dg))
    if Dout >= Din:
        return _linear(l, W, pr, True)            # σ.(W * x .+ b): GEMM with the bias/relu epilogue
    from .layers import _bias_act
    return _bias_act(l, pr)


# ---------------------------------------------------------------------------------------------------------
# bench.py --gpus N>1
# ---------------------------------------------------------------------------------------------------------
def bench_multi(args, world, rank, dev, seed, ClockSampler, measured_peaks, cpu_leg=None, parity=None):
    import gnnb200 as gnn
    n, E, D = args.nodes, args.edges, args.dim
    # NCCL opens its peer-to-peer connections lazily, on the first all-to-all (seconds at 8 ranks): communicator set-up is
    # not shard construction, so it is paid (and reported) before the plan timer starts
    torch.cuda.synchronize()
    tc0 = time.perf_counter()
    warm = torch.zeros(world * 4, device=dev)
    warm_out = torch.empty_like(warm)
    dist.all_to_all_single(warm_out, warm)
    dist.all_reduce(warm)
    dist.all_gather([torch.empty_like(warm) for _ in range(world)], warm)
    torch.cuda.synchronize()
    dist.barrier()
    t_comm = time.perf_counter() - tc0
    del warm, warm_out
    t0 = time.perf_counter()
    # every rank generates the counter-based edge list chunk by chunk and keeps its shard (csrc/shard.cu); 'balanced'
    # ownership deals the nodes to the ranks by decreasing degree (one extra pass over the generated chunks)
    ownership = os.environ.get("GNNB_OWNERSHIP", "balanced")
    dg = DistGraph.from_rmat(n, E, seed, device=dev, add_self_loops=True, ownership=ownership,
                             chunk_edges=int(os.environ.get("GNNB_CHUNK_EDGES", str(1 << 26))))
    torch.cuda.synchronize()
    dg.timing["constructor_total_ms"] = (time.perf_counter() - t0) * 1e3
    tg0 = time.perf_counter()
    dg.gcn_c()
    torch.cuda.synchronize()
    dg.timing["gcn_norm_and_its_halo_ms"] = (time.perf_counter() - tg0) * 1e3
    t_plan = time.perf_counter() - t0
    torch.cuda.empty_cache()

    torch.manual_seed(0)
    layer = gnn.GCNConv(D, D, torch.relu, device=dev)   # same seed => same weights on every rank
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = gnn.unrows(torch.randn(dg.n_local, D, device=dev, generator=gen)).requires_grad_(True)
    dy = gnn.unrows(torch.randn(dg.n_local, D, device=dev, generator=gen))

    def step():
        x.grad = None
        layer.weight.grad = None
        layer.bias.grad = None
        y = dist_gcn_conv(layer, dg, x)
        y.backward(dy)
        dist.all_reduce(layer.weight.grad)
        dist.all_reduce(layer.bias.grad)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    l0 = gnn.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clocks:
        torch.cuda.synchronize()
        dist.barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    launches = gnn.launch_count() - l0

    # the fused kernel alone on this rank's forward shard (CUDA events on the launch stream), max over ranks
    c, cf, cb = dg.gcn_c()
    xr = gnn.rows(x.detach())
    x.grad = None
    layer.weight.grad = None
    torch.cuda.empty_cache()
    hp = dg.halo_ptr(dg.fwd, xr)                        # the halo rows where the step's own exchange puts them
    out = torch.empty_like(xr)
    st = torch.cuda.current_stream(dev).cuda_stream

    def kern():
        _lib.check(lib.gnnb_propagate_halo(dg.fwd.plan.h, _lib.COPY_XJ, _lib.SUM, xr.data_ptr(), hp if dg.fwd.n_halo else None,
                                           dg.n_local, None, cf.data_ptr(), c.data_ptr(), D, out.data_ptr(), st))

    for _ in range(3):
        kern()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record(); kern(); b.record()
    torch.cuda.synchronize()
    kms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    # halo exchange alone
    for _ in range(2):
        dg.halo_ptr(dg.fwd, xr)
    torch.cuda.synchronize(); dist.barrier()
    ev0.record()
    for _ in range(args.steps):
        dg.halo_ptr(dg.fwd, xr)
    ev1.record()
    torch.cuda.synchronize()
    hms = ev0.elapsed_time(ev1) / args.steps
    Es = dg.fwd.num_edges
    alg = Es * (4 * D + 4) + 4 * (dg.n_local + 1) + 4 * D * dg.n_local
    stats = torch.tensor([kms, hms, float(Es), float(dg.n_local), float(dg.fwd.n_halo), float(dg.bwd.n_halo), alg / (kms * 1e-3) / 1e9],
                         device=dev, dtype=torch.float64)
    allst = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(allst, stats)
    allst = torch.stack(allst).cpu()
    peak, peak_src = measured_peaks()

    e2e = None
    if not args.no_e2e:
        xh = torch.empty(dg.n_local, D, pin_memory=True).normal_()
        dyh = torch.empty(dg.n_local, D, pin_memory=True).normal_()
        yh = torch.empty(dg.n_local, D, pin_memory=True)
        dxh = torch.empty(dg.n_local, D, pin_memory=True)

        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)

        def step_host():
            # uploads on s_in, downloads on s_out (PCIe is full duplex), compute on the main stream
            with torch.cuda.stream(s_in):
                xd_raw = xh.to(dev, non_blocking=True)
                ev_x = torch.cuda.Event(); ev_x.record(s_in)
                dyd_raw = dyh.to(dev, non_blocking=True)
                ev_dy = torch.cuda.Event(); ev_dy.record(s_in)
            main.wait_event(ev_x)
            xd = gnn.unrows(xd_raw).requires_grad_(True)
            layer.weight.grad = None
            layer.bias.grad = None
            y = dist_gcn_conv(layer, dg, xd)
            s_out.wait_stream(main)
            with torch.cuda.stream(s_out):
                yh.copy_(gnn.rows(y.detach()), non_blocking=True)
            main.wait_event(ev_dy)
            y.backward(gnn.unrows(dyd_raw))
            s_out.wait_stream(main)
            with torch.cuda.stream(s_out):
                dxh.copy_(gnn.rows(xd.grad), non_blocking=True)
            dist.all_reduce(layer.weight.grad)
            wg = layer.weight.grad.cpu()
            main.wait_stream(s_out)
            xd_raw.record_stream(main); dyd_raw.record_stream(main)
            return wg

        ke = max(2, min(args.steps, 5))
        step_host()
        torch.cuda.synchronize(); dist.barrier()
        ev0.record()
        for _ in range(ke):
            step_host()
        ev1.record()
        torch.cuda.synchronize()
        ems = torch.tensor([ev0.elapsed_time(ev1) / ke], device=dev)
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems = float(ems)
        e2e = {"value": E / (ems * 1e-3), "unit": "edges/s", "ms_per_step": ems, "steps": ke,
               "h2d_bytes_per_step": 2 * 4 * n * D, "d2h_bytes_per_step": 2 * 4 * n * D + 4 * D * D * world,
               "api": "gnnb200.partition.dist_gcn_conv on pinned host slices (all ranks; bytes are whole-job)"}

    par = None
    if parity is not None:                                   # all ranks: the checker runs collectives
        del x, dy, xr, out
        layer.weight.grad = None
        torch.cuda.empty_cache()
        par = parity(dg, layer)
    cpu = None
    if rank == 0 and cpu_leg is not None:
        cpu = cpu_leg()                                      # the oracle port on the bounded sample, host cores of rank 0
    if rank == 0:
        worst = int(torch.argmax(allst[:, 0]))
        line = {
            "metric": "edges/sec fwd+bwd GCNConv 128-dim on 100M-edge graph" if D == 128 else f"edges/sec fwd+bwd GCNConv {D}-dim (RMAT N={n} E={E})",
            "value": E / (ms * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GCNConv {D}->{D} (add_self_loops, relu, bias) fwd+bwd on RMAT N={n} E={E} seed {seed} "
                                   f"(BASELINE configs[{1 if D == 128 else 4}]), node-partitioned over {world} GPUs, halo exchange over NVLink",
                       "parallelism": f"node-partition x{world}, {ownership} ownership, shards built on the device from generated chunks",
                       "halo_exchange": os.environ.get("GNNB_HALO", "push") + (" (one kernel writes rows into peer halo buffers over NVLink, CUDA IPC)" if os.environ.get("GNNB_HALO", "push") == "push" else " (pack kernel + NCCL all_to_all_single)"),
                       "l2": "per-GPU features and halo buffers are far larger than the 126 MB L2",
                       "plan_build_ms": t_plan * 1e3, "plan_build_phases_ms_rank0": {k: round(v, 1) for k, v in dg.timing.items()},
                       "nccl_connection_setup_ms": t_comm * 1e3, "chunk_edges": 128,
                       "per_rank": {"kernel_ms": allst[:, 0].tolist(), "halo_exchange_ms": allst[:, 1].tolist(),
                                    "shard_edges": allst[:, 2].tolist(), "n_local": allst[:, 3].tolist(),
                                    "halo_rows_fwd": allst[:, 4].tolist(), "halo_rows_bwd": allst[:, 5].tolist()}},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": float(allst[worst, 6]), "peak": peak, "unit": "GB/s",
                         "frac": float(allst[worst, 6]) / peak, "traffic": None,
                         "kernel": "gnnb::seg_lean_kernel (sum, per-edge scale stream, halo base) over [local|halo] rows (slowest rank)",
                         "peak_source": peak_src,
                         "halo": {"bytes_received_slowest_rank": float(allst[:, 4].max()) * D * 4,
                                  "ms": float(allst[:, 1].max()),
                                  "GBps_per_gpu": float(allst[:, 4].max()) * D * 4 / (float(allst[:, 1].max()) * 1e-3) / 1e9,
                                  "nvlink_peak_GBps": 770.0}},
            "cpu_baseline": cpu, "parity_rel_err": par,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
