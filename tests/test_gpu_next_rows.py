"""SURVEY.md §8f rows built on the same kernels: graph-level readout (GNNlib/test/utils.jl:13-56 transcribed) and the
layers that re-parameterise propagate (graph_conv, gin_conv, sgc_conv, agnn_conv) against oracle compositions."""
import operator

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _batched(gnn, rng, ngraphs=5, n=10, e=60, Dx=2, De=3):
    gs = []
    for _ in range(ngraphs):
        s = rng.integers(1, n + 1, e)
        t = rng.integers(1, n + 1, e)
        gs.append(gnn.GNNGraph(s, t, num_nodes=n, ndata={"x": gnn.colmajor(torch.rand(Dx, n))},
                               edata={"e": gnn.colmajor(torch.rand(De, e))}))
    return gnn.batch(gs).cuda(), gs


def test_reduce_softmax_broadcast(gnn):
    rng = np.random.default_rng(0)
    g, gs = _batched(gnn, rng)
    x, e = g.ndata["x"], g.edata["e"]
    r = gnn.reduce_nodes(gnn.mean, g, x)
    assert r.shape == (2, g.num_graphs)
    assert torch.allclose(r[:, 1].cpu(), gs[1].ndata["x"].mean(dim=1), rtol=1e-6)           # utils.jl:16-17
    assert torch.equal(gnn.reduce_nodes(gnn.mean, gnn.graph_indicator(g), x), r)             # utils.jl:19-20
    r = gnn.reduce_edges(gnn.mean, g, e)
    assert r.shape == (3, g.num_graphs)
    assert torch.allclose(r[:, 1].cpu(), gs[1].edata["e"].mean(dim=1), rtol=1e-6)           # utils.jl:24-26
    for aggr, fn in ((operator.add, torch.sum), (max, torch.amax), (min, torch.amin)):
        assert torch.allclose(gnn.reduce_nodes(aggr, g, x)[:, 3].cpu(), fn(gs[3].ndata["x"], dim=1), rtol=1e-6)
    r = gnn.softmax_nodes(g, x)
    assert r.shape == x.shape
    assert torch.allclose(r[:, :10].cpu(), torch.softmax(gs[0].ndata["x"], dim=1), rtol=1e-5, atol=1e-7)  # utils.jl:30-32
    r = gnn.softmax_edges(g, e)
    assert r.shape == e.shape
    assert torch.allclose(r[:, :60].cpu(), torch.softmax(gs[0].edata["e"], dim=1), rtol=1e-5, atol=1e-6)  # utils.jl:36-38
    z = gnn.colmajor(torch.rand(4, g.num_graphs).cuda())
    r = gnn.broadcast_nodes(g, z)
    assert r.shape == (4, g.num_nodes)
    assert torch.equal(r[:, 0], z[:, 0]) and torch.equal(r[:, 9], z[:, 0]) and torch.equal(r[:, 10], z[:, 1])   # utils.jl:42-47
    r = gnn.broadcast_edges(g, z)
    assert r.shape == (4, g.num_edges)
    assert torch.equal(r[:, 0], z[:, 0]) and torch.equal(r[:, 59], z[:, 0]) and torch.equal(r[:, 60], z[:, 1])  # utils.jl:51-56
    # gradients flow through the library's pullbacks
    xg = x.clone().requires_grad_(True)
    (gnn.softmax_nodes(g, xg) * gnn.broadcast_nodes(g, gnn.reduce_nodes(max, g, xg))).sum().backward()
    assert torch.isfinite(xg.grad).all()

    class L:
        aggr = gnn.mean
        fgate = staticmethod(lambda v: v[:1])
        ffeat = staticmethod(lambda v: v)
    assert gnn.global_pool(L, g, x).shape == (2, 5)
    u = gnn.global_attention_pool(L, g, x)
    a = torch.softmax(gs[2].ndata["x"][:1], dim=1)
    assert torch.allclose(u[:, 2].cpu(), (a * gs[2].ndata["x"]).sum(dim=1), rtol=1e-5)


class _NT:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def test_reparameterised_layers(gnn, oracle):
    rng = np.random.default_rng(1)
    n, E, Din, Dout = 300, 2000, 16, 8
    s = rng.integers(1, n + 1, E); t = rng.integers(1, n + 1, E)
    g = gnn.GNNGraph(s, t, num_nodes=n).cuda()
    x = rng.standard_normal((n, Din)).astype(np.float32)
    xt = gnn.unrows(torch.as_tensor(x).cuda())
    x64 = x.astype(np.float64)
    dev = lambda a: torch.as_tensor(a.astype(np.float32)).cuda()
    W1 = rng.standard_normal((Dout, Din)) / 4; W2 = rng.standard_normal((Dout, Din)) / 4; b = rng.standard_normal(Dout)
    # graph_conv (conv.jl:102-108)
    l = _NT(weight1=dev(W1), weight2=dev(W2), bias=dev(b), σ=torch.relu, aggr=gnn.mean)
    m = oracle.propagate_unfused("mean", s, t, n, x64)
    ref = np.maximum(x64 @ W1.T + m @ W2.T + b, 0)
    assert rel(gnn.rows(gnn.graph_conv(l, g, xt)).cpu(), ref) < 5e-6
    # gin_conv (conv.jl:250-256) with nn = identity-ish linear map
    l = _NT(nn=lambda v: 2.0 * v, ϵ=0.25, aggr=operator.add)
    ref = 2.0 * (1.25 * x64 + oracle.propagate_unfused("+", s, t, n, x64))
    assert rel(gnn.rows(gnn.gin_conv(l, g, xt)).cpu(), ref) < 5e-6
    # sgc_conv (conv.jl:407-448), k = 2, self loops
    W = rng.standard_normal((Dout, Din)) / 4
    l = _NT(weight=dev(W), bias=dev(b), k=2, add_self_loops=True, use_edge_weight=False)
    s2, t2 = oracle.add_self_loops(s, t, n)
    h = x64 @ W.T
    for _ in range(2):
        h, _c = oracle.gcn_propagate(s2, t2, n, h)
    assert rel(gnn.rows(gnn.sgc_conv(l, g, xt)).cpu(), h + b) < 5e-6
    # agnn_conv (conv.jl:337-352)
    l = _NT(add_self_loops=True, β=torch.tensor(1.5, device="cuda"))
    xn = x64 / np.sqrt((x64 ** 2).sum(1, keepdims=True))
    cos = (xn[t2 - 1] * xn[s2 - 1]).sum(1, keepdims=True)
    alpha = oracle.softmax_edge_neighbors(t2, n, 1.5 * cos)
    ref = oracle.scatter("+", alpha * x64[s2 - 1], t2, n)
    assert rel(gnn.rows(gnn.agnn_conv(l, g, xt)).cpu(), ref) < 1e-5


# ------------------------------------------------------------------ node-partition shards built on the device (csrc/shard.cu)
@pytest.mark.parametrize("world,ownership", [(2, "contiguous"), (3, "cyclic"), (4, "balanced")])
def test_shard_builder_matches_the_torch_restatement(gnn, oracle, world, ownership):
    """gnnb_shard_builder_* (chunked, stable compaction, sort + unique halo, renaming, self loops, plan) for every rank of a
    `world`, on one GPU, against partition.build_shard (torch on CPU): halo lists, request counts and the plan's CSR
    (rowptr, col, eid) must be identical integers.  Needs no process group."""
    import ctypes as C
    from gnnb200 import partition as P
    lib = gnn._lib.lib
    rng = np.random.default_rng(world)
    n, E = 997, 20000
    s = np.minimum((rng.random(E) ** 3 * n).astype(np.int64), n - 1) + 1          # skewed: low ids are hubs
    t = np.minimum((rng.random(E) ** 2 * n).astype(np.int64), n - 1) + 1
    sd, td = torch.as_tensor(s).cuda(), torch.as_tensor(t).cuda()
    s0, t0 = torch.as_tensor(s) - 1, torch.as_tensor(t) - 1
    relabel = relabel_dev = None
    if ownership == "balanced":
        cost = torch.zeros(n, dtype=torch.int32, device="cuda")
        relabel_dev = torch.empty(n, dtype=torch.int32, device="cuda")
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        gnn._lib.check(lib.gnnb_degree_accumulate(sd.data_ptr(), td.data_ptr(), E, 8, 1, n, cost.data_ptr(), None))
        gnn._lib.check(lib.gnnb_balanced_relabel(cost.data_ptr(), n, world, relabel_dev.data_ptr(), order.data_ptr(), None))
        deg = torch.bincount(s0, minlength=n) + torch.bincount(t0, minlength=n)
        assert torch.equal(cost.cpu().long(), deg)
        by_degree = torch.sort(deg, descending=True, stable=True).indices
        pos = torch.arange(n)
        r, j = pos // world, pos % world
        o = torch.where((r % 2 == 1) & (r < n // world), world - 1 - j, j)
        relabel = torch.empty(n, dtype=torch.int32)
        relabel[by_degree] = (r * world + o).to(torch.int32)
        assert torch.equal(relabel_dev.cpu(), relabel)
        assert torch.equal(order.cpu().long()[relabel.long()], pos)
    bounds = [0, 100, 400, n][:world] + [n] if ownership == "contiguous" else None
    first = P.ownership_first(n, world, ownership, bounds)
    ps, pt = P.to_pid(s0, world, first, ownership, relabel), P.to_pid(t0, world, first, ownership, relabel)
    for rank in range(world):
        b = C.c_void_p()
        barr = (C.c_int64 * (world + 1))(*first) if ownership == "contiguous" else None
        gnn._lib.check(lib.gnnb_shard_builder_create(C.byref(b), n, world, rank, 0 if ownership == "contiguous" else 1, barr,
                                                     None if relabel_dev is None else relabel_dev.data_ptr()))
        try:
            for a in range(0, E, 7001):                                               # ragged chunks
                gnn._lib.check(lib.gnnb_shard_builder_add(b, sd[a:a + 7001].data_ptr(), td[a:a + 7001].data_ptr(),
                                                          min(7001, E - a), 8, 1, None))
            for direction, (key0, other0) in enumerate(((pt, ps), (ps, pt))):
                ref = P.build_shard(key0, other0, first[rank], first[rank + 1], first, True)
                h = C.c_void_p()
                nl, nh, ne = C.c_int64(), C.c_int64(), C.c_int64()
                rc = (C.c_int64 * world)()
                gnn._lib.check(lib.gnnb_shard_builder_finish(b, direction, 1, C.byref(h), C.byref(nl), C.byref(nh), C.byref(ne), rc, None))
                plan = gnn.graph._Plan(h.value, torch.device("cuda"))
                assert (nl.value, nh.value, ne.value) == (ref["n_local"], ref["halo"].numel(), ref["row"].numel())
                assert list(rc) == ref["recv_counts"]
                hl = torch.empty(max(nh.value, 1), dtype=torch.int32, device="cuda")
                gnn._lib.check(lib.gnnb_shard_builder_halo(b, direction, hl.data_ptr(), None))
                assert torch.equal(hl[:nh.value].cpu(), ref["halo_local"])
                rowptr, col, eid = np.empty(nl.value + 1, np.int32), np.empty(ne.value, np.int32), np.empty(ne.value, np.int32)
                gnn._lib.check(lib.gnnb_graph_csr(plan.h, 0, rowptr.ctypes.data, col.ctypes.data, eid.ctypes.data, None))
                rp_ref, col_ref, perm_ref = oracle.csr(ref["row"].numpy() + 1, ref["col"].numpy() + 1, nl.value)
                assert np.array_equal(rowptr, rp_ref) and np.array_equal(col, col_ref) and np.array_equal(eid, perm_ref)
        finally:
            lib.gnnb_shard_builder_destroy(b)


@pytest.mark.parametrize("relu_flag,with_bias,bwd", [(1, True, True), (0, False, True), (1, True, False)])
def test_gcn_conv_step_host_entry(gnn, oracle, relu_flag, with_bias, bwd):
    """gnnb_gcn_conv_step_host — a whole GCNConv forward (+ backward) on HOST arrays through the C ABI, the call bench.py times
    as e2e — against the oracle composition in fp64 (forward) and against the device-side layer (bit for bit: same kernels)."""
    lib = gnn._lib.lib
    n, E, D = 3000, 40000, 128
    rng = np.random.default_rng(7)
    s, t = oracle.rmat(n, E, 17)
    g = gnn.GNNGraph(s, t, num_nodes=n).cuda()
    g2 = gnn.add_self_loops(g)
    x = rng.standard_normal((n, D)).astype(np.float32)
    dy = rng.standard_normal((n, D)).astype(np.float32)
    W = (rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32)
    b = rng.standard_normal(D).astype(np.float32)
    y, dx, dW, db = np.empty_like(x), np.empty_like(x), np.empty_like(W), np.empty_like(b)
    ptr = lambda a: a.ctypes.data
    gnn._lib.check(lib.gnnb_gcn_conv_step_host(g2.plan().h, ptr(x), ptr(W), ptr(b) if with_bias else None, relu_flag, D, D,
                                               ptr(dy) if bwd else None, ptr(y), ptr(dx) if bwd else None,
                                               ptr(dW) if bwd else None, ptr(db) if (bwd and with_bias) else None))
    s2, t2 = oracle.add_self_loops(s, t, n)
    p, c = oracle.gcn_propagate(s2, t2, n, x.astype(np.float64))
    pre = p @ W.astype(np.float64).T + (b if with_bias else 0)
    ref = np.maximum(pre, 0) if relu_flag else pre
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-5
    # the device-side layer runs the same kernels: identical bits
    layer = gnn.GCNConv(D, D, torch.relu if relu_flag else gnn.layers.identity, bias=with_bias, device="cuda")
    with torch.no_grad():
        layer.weight.copy_(torch.as_tensor(W))
        if with_bias:
            layer.bias.copy_(torch.as_tensor(b))
    xt = gnn.unrows(torch.as_tensor(x).cuda()).requires_grad_(True)
    yt = layer(g, xt)
    assert np.array_equal(gnn.rows(yt.detach()).cpu().numpy(), y)
    if bwd:
        yt.backward(gnn.unrows(torch.as_tensor(dy).cuda()))
        assert np.array_equal(gnn.rows(xt.grad).cpu().numpy(), dx)
        assert np.array_equal(layer.weight.grad.cpu().numpy(), dW)
        if with_bias:
            assert np.array_equal(layer.bias.grad.cpu().numpy(), db)
