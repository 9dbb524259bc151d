"""Neighbour sampling (graphneuralnetworks.jl_b200/sampling.py over csrc/sample.cu).

The reference draws from Julia's RNG, so these are the size-independent properties of GNNGraphs/src/sampling.jl:68-121
and GNNGraphs/test/sampling.jl (every sampled edge is an in/out-edge of its node, counts = min(K, deg) or K, no
repeats without replacement, EID / NID bookkeeping, induced subgraphs), plus a uniformity check of the subset sampler.

Back ends: the CPU test double and (under -m gpu) the CUDA kernels of csrc/sample.cu.
"""
import numpy as np
import pytest
import torch


@pytest.fixture
def bs(be):
    return be


def idx(a, dev):
    return torch.as_tensor(np.asarray(a), dtype=torch.int64).to(dev)


def graph(gnn, rng, n, E, dev, **kw):
    s = rng.integers(1, n + 1, E); t = rng.integers(1, n + 1, E)
    return gnn.GNNGraph(idx(s, dev), idx(t, dev), num_nodes=n, **kw), s, t


@pytest.mark.parametrize("direction", ["in", "out"])
@pytest.mark.parametrize("K,replace", [(-1, False), (3, False), (50, False), (4, True), (-1, True)])
def test_sample_edge_ids_properties(gnn, bs, direction, K, replace):
    rng = np.random.default_rng(0)
    n, E = 40, 600
    g, s, t = graph(gnn, rng, n, E, bs.dev)
    key = t if direction == "in" else s
    key[key == 5] = 6                                                   # node 5 has no edges in this direction
    g = gnn.GNNGraph(idx(s, bs.dev), idx(t, bs.dev), num_nodes=n)
    deg = np.bincount(key, minlength=n + 1)
    nodes = np.array([2, 5, 9, 9, 40, 1])
    eids, offsets = gnn.sample_edge_ids(g, idx(nodes, bs.dev), K, dir=direction, replace=replace, seed=11)
    eids, offsets = eids.cpu().numpy(), offsets.cpu().numpy()
    assert offsets[0] == 0 and offsets[-1] == len(eids)
    for j, v in enumerate(nodes):
        mine = eids[offsets[j]:offsets[j + 1]]
        d = deg[v]
        want = 0 if d == 0 else ((K if K > 0 else d) if replace else (min(K, d) if K > 0 else d))
        assert len(mine) == want
        assert (key[mine - 1] == v).all()                               # 1-based COO ids of edges incident to v
        if not replace:
            assert len(set(mine.tolist())) == len(mine)
        if not replace and want == d:
            assert sorted(mine.tolist()) == (np.nonzero(key == v)[0] + 1).tolist()
    again, _ = gnn.sample_edge_ids(g, idx(nodes, bs.dev), K, dir=direction, replace=replace, seed=11)
    assert np.array_equal(again.cpu().numpy(), eids)                    # same seed, same draw
    with pytest.raises(AssertionError):
        gnn.sample_edge_ids(g, idx([1, n + 1], bs.dev), K, dir=direction, replace=replace)
    e, o = gnn.sample_edge_ids(g, idx([], bs.dev), K, dir=direction, replace=replace)
    assert e.numel() == 0 and o.tolist() == [0]


def test_subsets_are_uniform(gnn, bs):
    """one node with 6 in-edges, K=2: each of the 15 pairs should come up ~1/15 of the time over many seeds"""
    s = np.arange(1, 7); t = np.full(6, 7)
    g = gnn.GNNGraph(idx(s, bs.dev), idx(t, bs.dev), num_nodes=7)
    counts = {}
    trials = 3000
    nodes = idx([7] * 10, bs.dev)                                       # ten independent draws per call
    for seed in range(trials // 10):
        eids, off = gnn.sample_edge_ids(g, nodes, 2, seed=seed)
        for pair in eids.cpu().numpy().reshape(-1, 2):
            key = tuple(sorted(pair.tolist()))
            counts[key] = counts.get(key, 0) + 1
    assert len(counts) == 15
    exp = trials / 15
    chi2 = sum((c - exp) ** 2 / exp for c in counts.values())
    assert chi2 < 45                                                    # 14 dof: P(chi2 > 45) ~ 4e-5


def test_sample_neighbors_graph(gnn, bs):
    """the docstring example of GNNGraphs/src/sampling.jl:24-66, as properties"""
    rng = np.random.default_rng(1)
    n, E = 20, 100
    e = torch.arange(1, E + 1, dtype=torch.float32).to(bs.dev)
    x = gnn.colmajor(torch.arange(n, dtype=torch.float32).repeat(2, 1).to(bs.dev))
    g, s, t = graph(gnn, rng, n, E, bs.dev, edata={"e": e}, ndata={"x": x})
    nodes = [2, 3]
    sg = gnn.sample_neighbors(g, idx(nodes, bs.dev))                    # all in-edges of nodes 2 and 3
    m = int(np.isin(t, nodes).sum())
    assert sg.num_nodes == n and sg.num_edges == m and sg.edata["EID"].shape == (m,)
    eid = sg.edata["EID"].cpu().numpy()
    assert np.array_equal(sg.s.cpu().numpy(), s[eid - 1]) and np.array_equal(sg.t.cpu().numpy(), t[eid - 1])
    assert torch.equal(sg.edata["e"].cpu(), torch.as_tensor(eid, dtype=torch.float32))
    sg = gnn.sample_neighbors(g, idx(nodes, bs.dev), 5, replace=True, seed=3)
    assert sg.num_edges == 10 and np.isin(sg.t.cpu().numpy(), nodes).all()
    sg = gnn.sample_neighbors(g, idx(nodes, bs.dev), dropnodes=True)
    nid = sg.ndata["NID"].cpu().numpy()
    assert nid[:2].tolist() == nodes and len(set(nid.tolist())) == len(nid) == sg.num_nodes
    assert set(nid.tolist()) == set(nodes) | set(s[np.isin(t, nodes)].tolist())
    eid = sg.edata["EID"].cpu().numpy()
    assert np.array_equal(nid[sg.s.cpu().numpy() - 1], s[eid - 1]) and np.array_equal(nid[sg.t.cpu().numpy() - 1], t[eid - 1])
    assert torch.equal(sg.ndata["x"][0].cpu(), torch.as_tensor(nid - 1, dtype=torch.float32))
    # extra nodes appear in order of first appearance among the sampled sources (Julia's setdiff)
    extra, seen = [], set(nodes)
    for v in s[eid - 1]:
        if v not in seen:
            seen.add(v); extra.append(v)
    assert nid[2:].tolist() == extra
    sg = gnn.sample_neighbors(g, idx(nodes, bs.dev), 2, dir="out", dropnodes=True, seed=4)
    assert np.isin(sg.ndata["NID"].cpu().numpy()[sg.s.cpu().numpy() - 1], nodes).all()


def test_induced_subgraph(gnn, bs):
    """GNNGraphs/src/sampling.jl:129-170 docstring: s=[1,2], t=[2,3], nodes [1,2] -> 2 nodes, 1 edge"""
    x = gnn.colmajor(torch.rand(32, 3).to(bs.dev))
    g = gnn.GNNGraph(idx([1, 2], bs.dev), idx([2, 3], bs.dev), ndata={"x": x, "y": torch.rand(3).to(bs.dev)},
                     edata={"e": torch.tensor([5.0, 7.0]).to(bs.dev)})
    sub = gnn.induced_subgraph(g, idx([1, 2], bs.dev))
    assert sub.num_nodes == 2 and sub.num_edges == 1 and sub.s.tolist() == [1] and sub.t.tolist() == [2]
    assert sub.ndata["x"].shape == (32, 2) and sub.ndata["y"].shape == (2,) and sub.edata["e"].tolist() == [5.0]
    assert torch.equal(sub.ndata["x"].cpu(), x[:, :2].cpu())
    sub = gnn.induced_subgraph(g, idx([3, 1], bs.dev))                  # no edge between 1 and 3: isolated nodes
    assert sub.num_nodes == 2 and sub.num_edges == 0
    assert gnn.induced_subgraph(g, idx([], bs.dev)).num_nodes == 0
    rng = np.random.default_rng(2)
    g, s, t = graph(gnn, rng, 30, 300, bs.dev)
    nodes = np.array([7, 3, 19, 22, 4])
    sub = gnn.induced_subgraph(g, idx(nodes, bs.dev))
    keep = np.isin(s, nodes) & np.isin(t, nodes)
    assert sub.num_edges == int(keep.sum())
    pairs = sorted(zip(nodes[sub.s.cpu().numpy() - 1].tolist(), nodes[sub.t.cpu().numpy() - 1].tolist()))
    assert pairs == sorted(zip(s[keep].tolist(), t[keep].tolist()))


def test_neighbor_loader(gnn, bs):
    """GNNGraphs/test/samplers.jl (commented out upstream) as properties: batch count, inputs present, sizes bounded"""
    rng = np.random.default_rng(3)
    n, E = 60, 400
    x = gnn.colmajor(torch.arange(n, dtype=torch.float32).reshape(1, n).to(bs.dev))
    g, s, t = graph(gnn, rng, n, E, bs.dev, ndata={"x": x})
    loader = gnn.NeighborLoader(g, num_neighbors=[2, 2], input_nodes=[1, 2, 3, 4, 5], num_layers=2, batch_size=2, seed=9)
    batches = list(loader)
    assert len(batches) == len(loader) == 3
    for b, mb in enumerate(batches):
        inputs = [1, 2, 3, 4, 5][2 * b:2 * b + 2]
        ids = (mb.ndata["x"][0].cpu().numpy() + 1).astype(int)          # original ids of the mini-batch nodes
        assert ids[:len(inputs)].tolist() == inputs
        assert len(set(ids.tolist())) == mb.num_nodes <= len(inputs) * (1 + 2 + 4)
        keep = np.isin(s, ids) & np.isin(t, ids)
        assert mb.num_edges == int(keep.sum())                          # induced: every edge among the reached nodes
    one = list(gnn.NeighborLoader(g, num_neighbors=[2], input_nodes=[1, 2], num_layers=1, batch_size=10, seed=1))
    assert len(one) == 1 and one[0].num_nodes <= 2 * 3
    zero = list(gnn.NeighborLoader(g, num_neighbors=[0], input_nodes=[1, 2], num_layers=1, batch_size=2))
    assert zero[0].num_nodes == 2


# ---------------------------------------------------------------- the device sampler's own code, compiled for the host
def _positions(gnn, deg, K, replace, seed, j):
    import ctypes as C
    out = np.empty(max(deg, K, 1), dtype=np.int64)
    k = C.c_int64(0)
    gnn._lib.check(gnn._lib.lib.gnnb_sample_positions_host(deg, K, int(replace), seed, j, out.ctypes.data, len(out),
                                                           C.byref(k)))
    return out[:k.value]


@pytest.mark.parametrize("deg,K", [(6, 2), (10, 3), (7, 5), (40, 3), (5, 4)])
def test_device_sampler_on_host_is_valid_and_uniform(gnn, deg, K):
    """gnnb_sample_positions_host runs sample_positions() of csrc/sample.cu — the function the kernel calls — on the CPU.
    (deg, K) choose Floyd's algorithm (K*K <= 4*deg) or selection sampling; both must give each position the same
    inclusion probability K/deg, and each pair the same joint probability, over the (seed, j) counter."""
    trials = 6000
    hits = np.zeros(deg)
    pair = np.zeros((deg, deg))
    for q in range(trials):
        p = _positions(gnn, deg, K, False, 1234 + q // 50, q % 50)
        assert len(p) == K and len(set(p.tolist())) == K and p.min() >= 0 and p.max() < deg
        hits[p] += 1
        pair[np.ix_(p, p)] += 1
    exp = trials * K / deg
    assert np.abs(hits - exp).max() < 5 * np.sqrt(exp)                  # 5 sigma on each marginal
    pe = trials * K * (K - 1) / (deg * (deg - 1))
    off = pair[~np.eye(deg, dtype=bool)]
    assert np.abs(off - pe).max() < 6 * np.sqrt(pe) + 3
    assert np.array_equal(_positions(gnn, deg, K, False, 7, 3), _positions(gnn, deg, K, False, 7, 3))
    assert not np.array_equal(_positions(gnn, 1000, 20, False, 7, 3), _positions(gnn, 1000, 20, False, 7, 4))


def test_device_sampler_on_host_edge_cases(gnn):
    assert _positions(gnn, 0, 3, False, 1, 0).size == 0 and _positions(gnn, 0, 3, True, 1, 0).size == 0
    assert _positions(gnn, 4, -1, False, 1, 0).tolist() == [0, 1, 2, 3]        # K <= 0: everything, adjacency order
    assert _positions(gnn, 4, 9, False, 1, 0).tolist() == [0, 1, 2, 3]         # K >= deg: everything
    p = _positions(gnn, 3, 8, True, 5, 2)                                     # with replacement: K draws, repeats allowed
    assert len(p) == 8 and p.min() >= 0 and p.max() < 3
    d = np.bincount(np.concatenate([_positions(gnn, 5, 10, True, s, 0) for s in range(800)]), minlength=5)
    assert np.abs(d - 1600).max() < 5 * np.sqrt(1600)
    s_path = _positions(gnn, 7, 5, False, 9, 1)                               # selection sampling keeps adjacency order
    assert s_path.tolist() == sorted(s_path.tolist())
