"""Edge-list transforms (graphneuralnetworks.jl_b200/transform.py over csrc/transform.cu) on both
back ends of the `be` fixture (CPU test double of the ABI / the CUDA library).  Integer results are compared with `==`.

Transcribed from the reference: GNNGraphs/test/transform.jl:284-322 (remove_self_loops, remove_multi_edges), :379-396
(to_bidirected known answer = the docstring example at src/transform.jl:447-490), :56-80 (unbatch round trip),
GNNGraphs/test/gnngraph.jl:75 (sort_edge_index of an edge_index pair).
"""
import operator

import numpy as np
import pytest
import torch


def ref_sort(u, v):
    """sortperm(collect(zip(u, v))) — GNNGraphs/src/utils.jl:41-45; Python's sort is stable like Julia's"""
    p = sorted(range(len(u)), key=lambda k: (int(u[k]), int(v[k])))
    return np.asarray(p, dtype=np.int64)


def idx(a, dev, dtype=torch.int64):
    return torch.as_tensor(np.asarray(a), dtype=dtype).to(dev)


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("E,n", [(1, 1), (7, 3), (1000, 17), (5000, 100000)])
def test_sort_edge_index(gnn, be, E, n, dtype):
    rng = np.random.default_rng(E + n)
    u = rng.integers(1, n + 1, E)
    v = rng.integers(1, n + 1, E)
    p = ref_sort(u, v)
    us, vs, perm = gnn.sort_edge_index(idx(u, be.dev, dtype), idx(v, be.dev, dtype), return_perm=True)
    assert us.dtype == dtype and vs.dtype == dtype and us.device.type == be.dev.type
    assert np.array_equal(us.cpu().numpy(), u[p]) and np.array_equal(vs.cpu().numpy(), v[p])
    assert np.array_equal(perm.cpu().numpy(), p)                       # stable: equal pairs keep their COO order
    us2, vs2 = gnn.sort_edge_index((idx(u, be.dev, dtype), idx(v, be.dev, dtype)))       # tuple form, gnngraph.jl:75
    assert torch.equal(us2, us) and torch.equal(vs2, vs)
    # sort by target = the other way round (color_refinement's use, utils.jl:370)
    ts, ss = gnn.sort_edge_index(idx(v, be.dev, dtype), idx(u, be.dev, dtype))
    q = ref_sort(v, u)
    assert np.array_equal(ts.cpu().numpy(), v[q]) and np.array_equal(ss.cpu().numpy(), u[q])


def test_sort_edge_index_edge_cases(gnn, be):
    e = torch.empty(0, dtype=torch.int64, device=be.dev)
    us, vs = gnn.sort_edge_index(e, e)
    assert us.numel() == 0 and vs.numel() == 0
    z = idx([0, 2, 0, 1], be.dev)                                      # 0-based ids sort too
    us, vs = gnn.sort_edge_index(z, idx([3, 0, 1, 1], be.dev))
    assert us.tolist() == [0, 0, 1, 2] and vs.tolist() == [1, 3, 1, 0]
    with pytest.raises(AssertionError):                                # negative ids: GNNB_EINDEX
        gnn.sort_edge_index(idx([1, -2], be.dev), idx([1, 1], be.dev))
    with pytest.raises(AssertionError):
        gnn.sort_edge_index(idx([1, 2], be.dev), idx([1], be.dev))


def _rand_graph(gnn, rng, n, E, dev):
    """distinct random edges without self loops, like rand_graph(n, m) in the reference's tests"""
    pairs = [(a, b) for a in range(1, n + 1) for b in range(1, n + 1) if a != b]
    sel = rng.choice(len(pairs), E, replace=False)
    s = np.asarray([pairs[k][0] for k in sel]); t = np.asarray([pairs[k][1] for k in sel])
    return gnn.GNNGraph(idx(s, dev), idx(t, dev), num_nodes=n), s, t


def _add_edges(gnn, g, s_new, t_new, **kw):
    dev = g.s.device
    return gnn.GNNGraph(torch.cat([g.s, idx(s_new, dev)]), torch.cat([g.t, idx(t_new, dev)]), num_nodes=g.num_nodes, **kw)


def same_edge_set(gnn, g2, g):
    a, b = gnn.sort_edge_index(gnn.edge_index(g2)), gnn.sort_edge_index(gnn.edge_index(g))
    return torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_remove_self_loops(gnn, be):
    """GNNGraphs/test/transform.jl:284-301"""
    rng = np.random.default_rng(0)
    g, s, t = _rand_graph(gnn, rng, 10, 20, be.dev)
    g1 = _add_edges(gnn, g, np.arange(1, 6), np.arange(1, 6))
    assert g1.num_edges == g.num_edges + 5
    g2 = gnn.remove_self_loops(g1)
    assert g2.num_edges == g.num_edges and same_edge_set(gnn, g2, g)
    E1 = g1.num_edges
    g1 = _add_edges(gnn, g, np.arange(1, 6), np.arange(1, 6),
                    edata={"e1": gnn.colmajor(torch.ones(3, E1, device=be.dev)), "e2": 2 * torch.ones(E1, device=be.dev)})
    g1 = gnn.set_edge_weight(g1, 3 * torch.ones(E1, device=be.dev))
    g2 = gnn.remove_self_loops(g1)
    assert g2.num_edges == g.num_edges and same_edge_set(gnn, g2, g)
    assert g2.w.shape == (g2.num_edges,) and g2.edata["e1"].shape == (3, g2.num_edges)
    assert g2.edata["e2"].shape == (g2.num_edges,)


def test_remove_multi_edges(gnn, be):
    """GNNGraphs/test/transform.jl:303-322"""
    rng = np.random.default_rng(1)
    g, s, t = _rand_graph(gnn, rng, 10, 20, be.dev)
    g1 = _add_edges(gnn, g, s[:5], t[:5])
    assert g1.num_edges == g.num_edges + 5
    g2 = gnn.remove_multi_edges(g1, aggr=operator.add)
    assert g2.num_edges == g.num_edges and same_edge_set(gnn, g2, g)
    so, to = gnn.sort_edge_index(gnn.edge_index(g))
    assert torch.equal(g2.s, so) and torch.equal(g2.t, to)             # the result is in (s, t) order
    E1 = g1.num_edges
    g1 = _add_edges(gnn, g, s[:5], t[:5],                              # default aggregation is +
                    edata={"e1": gnn.colmajor(torch.ones(3, E1, device=be.dev)), "e2": 2 * torch.ones(E1, device=be.dev)})
    g1 = gnn.set_edge_weight(g1, 3 * torch.ones(E1, device=be.dev))
    g2 = gnn.remove_multi_edges(g1)
    assert g2.num_edges == g.num_edges and same_edge_set(gnn, g2, g)
    e1, e2, w2 = g2.edata["e1"].cpu(), g2.edata["e2"].cpu(), g2.w.cpu()
    assert sum(bool((e1[:, i] == 2).all()) for i in range(g2.num_edges)) == 5
    assert int((e2 == 4).sum()) == 5 and int((w2 == 6).sum()) == 5
    assert int((e2 == 2).sum()) == 15 and int((w2 == 3).sum()) == 15


@pytest.mark.parametrize("aggr", ["+", "mean", "max", "min"])
def test_remove_multi_edges_aggregations(gnn, be, aggr):
    """against the reference's statement sequence (transform.jl:157-190) in numpy"""
    rng = np.random.default_rng(2)
    n, E = 12, 300                                                     # 144 possible pairs: plenty of repeats
    s = rng.integers(1, n + 1, E); t = rng.integers(1, n + 1, E)
    w = rng.uniform(0.5, 2, E).astype(np.float32)
    e = rng.standard_normal((E, 3)).astype(np.float32)
    op = {"+": operator.add, "mean": gnn.mean, "max": max, "min": min}[aggr]
    g = gnn.GNNGraph(idx(s, be.dev), idx(t, be.dev), torch.as_tensor(w).to(be.dev), num_nodes=n,
                     edata={"e": gnn.unrows(torch.as_tensor(e).to(be.dev))})
    g2 = gnn.remove_multi_edges(g, aggr=op)
    enc = (s - 1) * n + t                                              # edge_encoding, utils.jl:189-192
    perm = np.argsort(enc, kind="stable")
    encs = enc[perm]
    mask = np.concatenate([[True], encs[1:] > encs[:-1]])
    seg = np.cumsum(mask) - 1
    nu = int(mask.sum())
    red = {"+": np.add, "mean": np.add, "max": np.maximum, "min": np.minimum}[aggr]
    init = {"+": 0.0, "mean": 0.0, "max": -np.inf, "min": np.inf}[aggr]
    wr = np.full(nu, init); er = np.full((nu, 3), init)
    red.at(wr, seg, w[perm].astype(np.float64)); red.at(er, seg, e[perm].astype(np.float64))
    if aggr == "mean":
        cnt = np.bincount(seg, minlength=nu)
        wr, er = wr / cnt, er / cnt[:, None]
    assert g2.num_edges == nu
    assert np.array_equal(g2.s.cpu().numpy(), s[perm][mask]) and np.array_equal(g2.t.cpu().numpy(), t[perm][mask])
    assert np.allclose(g2.w.cpu().numpy(), wr, rtol=2e-6, atol=0)
    assert np.allclose(gnn.rows(g2.edata["e"]).cpu().numpy(), er, rtol=2e-6, atol=1e-7)
    # idempotent, and a graph without repeats only gets sorted
    g3 = gnn.remove_multi_edges(g2, aggr=op)
    assert torch.equal(g3.s, g2.s) and torch.equal(g3.t, g2.t) and torch.equal(g3.w, g2.w)
    assert torch.equal(g3.edata["e"], g2.edata["e"])
    with pytest.raises(AssertionError):                                # out-of-range target: GNNB_EINDEX
        gnn.remove_multi_edges(gnn.GNNGraph(idx([1, 2], be.dev), idx([2, 5], be.dev), num_nodes=3))


def test_to_bidirected_known_answer(gnn, be):
    """GNNGraphs/test/transform.jl:379-396 = the docstring example, src/transform.jl:447-490"""
    g = gnn.GNNGraph(idx([1, 2, 3, 3, 4], be.dev), idx([2, 3, 4, 4, 4], be.dev),
                     torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0], device=be.dev),
                     edata={"e": torch.tensor([10.0, 20.0, 30.0, 40.0, 50.0], device=be.dev)})
    g2 = gnn.to_bidirected(g)
    assert g2.num_nodes == g.num_nodes and g2.num_edges == 7
    assert g2.s.tolist() == [1, 2, 2, 3, 3, 4, 4] and g2.t.tolist() == [2, 1, 3, 2, 4, 3, 4]
    assert g2.w.tolist() == [1, 1, 2, 2, 3.5, 3.5, 5]
    assert g2.edata["e"].tolist() == [10.0, 10.0, 20.0, 20.0, 35.0, 35.0, 50.0]
    # is_bidirected and no multi edges
    pairs = set(zip(g2.s.tolist(), g2.t.tolist()))
    assert len(pairs) == g2.num_edges and all((b, a) in pairs for a, b in pairs)


def test_batch_unbatch_roundtrip(gnn, be):
    """GNNGraphs/test/transform.jl:56-80"""
    rng = np.random.default_rng(3)
    n, c, ngraphs = 20, 3, 10
    gs = []
    for _ in range(ngraphs):
        g, s, t = _rand_graph(gnn, rng, n, c * n, be.dev)
        g.ndata["x"] = gnn.colmajor(torch.rand(2, n, device=be.dev))
        g.edata["e"] = gnn.colmajor(torch.rand(3, c * n, device=be.dev))
        gs.append(g)
    gall = gnn.batch(gs)
    gs2 = gnn.unbatch(gall)
    assert len(gs2) == ngraphs
    for a, b in ((gs2[0], gs[0]), (gs2[-1], gs[-1]), (gs2[4], gs[4])):
        assert a.num_nodes == b.num_nodes and a.num_edges == b.num_edges and a.num_graphs == 1
        assert torch.equal(a.s, b.s) and torch.equal(a.t, b.t)
        assert torch.equal(a.ndata["x"], b.ndata["x"]) and torch.equal(a.edata["e"], b.edata["e"])
    g1, _, _ = _rand_graph(gnn, rng, 10, 20, be.dev)
    g2, _, _ = _rand_graph(gnn, rng, 5, 10, be.dev)
    u = gnn.unbatch(gnn.batch([g1, g2]))
    assert [x.num_nodes for x in u] == [10, 5] and [x.num_edges for x in u] == [20, 10]
    assert gnn.unbatch(g1) == [g1]


def test_csr_api(gnn, be):
    rng = np.random.default_rng(4)
    n, E = 30, 200
    s = rng.integers(1, n + 1, E); t = rng.integers(1, n + 1, E)
    g = gnn.GNNGraph(idx(s, be.dev), idx(t, be.dev), num_nodes=n)
    for transposed, key, other in ((False, t, s), (True, s, t)):
        rowptr, col, eid = (a.cpu().numpy() for a in gnn.csr(g, transposed))
        order = np.argsort(key, kind="stable")
        assert rowptr.dtype == np.int32 and np.array_equal(rowptr, np.concatenate([[0], np.cumsum(np.bincount(key - 1, minlength=n))]))
        assert np.array_equal(eid, order) and np.array_equal(col, other[order] - 1)
    # the degrees are the row lengths (GNNGraphs/test/query.jl:49-58)
    rowptr = gnn.csr(g)[0]
    assert torch.equal((rowptr[1:] - rowptr[:-1]).float(), gnn.degree(g, dir="in").to(rowptr.device).float())
