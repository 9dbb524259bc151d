"""The reference's own layer test items (GraphNeuralNetworks/test/layers/conv.jl), transcribed for the layers of this
repo: constructor fields and trainable-parameter counts, output sizes on TEST_GRAPHS (test_module.jl:159-180: the
4-cycle and the graph with an isolated vertex; COO only — dense / sparse adjacency graphs are outside the path), and
gradients that exist and are finite (the exact gradient values are checked in tests/test_layers.py).
Both back ends of the `be` fixture."""
import operator

import pytest
import torch

D_IN, D_OUT = 3, 5                                                     # test_module.jl:56-57


def ref_graphs(gnn, dev):
    adj1 = torch.tensor([[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 0, 1], [1, 0, 1, 0]])
    adj2 = torch.tensor([[0, 0, 0, 1], [0, 0, 0, 0], [0, 0, 0, 1], [1, 0, 1, 0]])      # vertex 2 is isolated
    out = []
    for A in (adj1, adj2):
        g = gnn.GNNGraph(A).to(dev)
        g.ndata["x"] = gnn.colmajor(torch.rand(D_IN, 4, device=dev))
        out.append(g)
    return out


def ntrainable(layer):
    return sum(1 for p in layer.parameters() if p.requires_grad)


def check(gnn, layer, g, *inputs, size):
    xs = [t.clone().requires_grad_(True) for t in inputs]
    y = layer(g, *xs)
    assert tuple(y.shape) == size
    y.sum().backward()
    for t in xs:
        assert t.grad is not None and torch.isfinite(t.grad).all()
    for p in layer.parameters():
        assert p.grad is None or torch.isfinite(p.grad).all()


def test_gcnconv(gnn, be):                                              # conv.jl:7-66
    dev = be.dev
    for g in ref_graphs(gnn, dev):
        for kw in (dict(), dict(sigma=torch.tanh, bias=False), dict(sigma=torch.tanh, add_self_loops=False)):
            if kw.get("add_self_loops") is False and g.num_edges == 4:
                continue                                                # 1/sqrt(0) on the isolated vertex: NaN there too
            check(gnn, gnn.GCNConv(D_IN, D_OUT, device=dev, **kw), g, g.x, size=(D_OUT, g.num_nodes))
        l = gnn.GCNConv(D_IN, D_OUT, device=dev)
        w = torch.zeros(D_OUT, D_IN, device=dev)
        x = g.x
        assert torch.equal(l(g, x, conv_weight=w), torch.zeros(D_OUT, g.num_nodes, device=dev))     # conv.jl:55-61
    # edge weights and the closed form (conv.jl:30-44) — s, t, w of the test
    s, t = [2, 3, 1, 3, 1, 2], [1, 1, 2, 2, 3, 3]
    w = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], device=dev)
    g = gnn.GNNGraph(torch.tensor(s), torch.tensor(t)).to(dev)
    x = gnn.colmajor(torch.ones(1, 3, device=dev))
    l = gnn.GCNConv(1, 1, add_self_loops=False, use_edge_weight=True, device=dev)
    with torch.no_grad():
        l.weight.fill_(1.0)
    d = gnn.degree(g, dir="in", edge_weight=w)
    y = l(g, x, w)
    assert abs(float(y[0, 0].detach()) - float(w[0] / (d[0] * d[1]).sqrt() + w[1] / (d[0] * d[2]).sqrt())) < 1e-5
    assert abs(float(y[0, 1].detach()) - float(w[2] / (d[1] * d[0]).sqrt() + w[3] / (d[1] * d[2]).sqrt())) < 1e-5
    assert torch.allclose(y, l(g, x, w, norm_fn=lambda dd: 1.0 / torch.sqrt(dd)))                  # conv.jl:43
    ww = w.clone().requires_grad_(True)
    l(g, x, ww).sum().backward()                                                                  # conv.jl:50
    assert ww.grad.shape == (6,) and ww.grad.dtype == torch.float32


def test_graphconv_sage_gin(gnn, be):                                   # conv.jl:122-141, 277-288, 327-337
    dev = be.dev
    assert ntrainable(gnn.GraphConv(2, 3)) == 3 and ntrainable(gnn.GraphConv(2, 3, bias=False)) == 2
    assert gnn.SAGEConv(D_IN, D_OUT).aggr is gnn.mean
    for g in ref_graphs(gnn, dev):
        check(gnn, gnn.GraphConv(D_IN, D_OUT, device=dev), g, g.x, size=(D_OUT, 4))
        check(gnn, gnn.GraphConv(D_IN, D_OUT, torch.tanh, bias=False, aggr=gnn.mean, device=dev), g, g.x, size=(D_OUT, 4))
        check(gnn, gnn.SAGEConv(D_IN, D_OUT, device=dev), g, g.x, size=(D_OUT, 4))
        check(gnn, gnn.SAGEConv(D_IN, D_OUT, torch.tanh, bias=False, aggr=operator.add, device=dev), g, g.x, size=(D_OUT, 4))
        nn = torch.nn.Linear(D_IN, D_OUT, device=dev)
        gin = gnn.GINConv(lambda v: gnn.unrows(nn(gnn.rows(v))), 0.001, aggr=gnn.mean)
        check(gnn, gin, g, g.x, size=(D_OUT, 4))
        assert not any(p is gin.eps for p in gin.parameters())         # ϵ is not trainable (conv.jl:287)


@pytest.mark.parametrize("cls", ["GATConv", "GATv2Conv"])
def test_attention_layers(gnn, be, cls):                                # conv.jl:154-180, 194-220
    dev = be.dev
    L = getattr(gnn, cls)
    for heads in (1, 2):
        for concat in (True, False):
            l = L(D_IN, D_OUT, torch.tanh, heads=heads, concat=concat, dropout=0, device=dev)
            for g in ref_graphs(gnn, dev):
                check(gnn, l, g, g.x, size=((heads * D_OUT if concat else D_OUT), 4))
    if cls == "GATv2Conv":                                              # edge features (conv.jl:203-211)
        ein = 3
        l = gnn.GATv2Conv((D_IN, ein), D_OUT, add_self_loops=False, dropout=0, device=dev)
        for g in ref_graphs(gnn, dev):
            e = gnn.colmajor(torch.rand(ein, g.num_edges, device=dev))
            check(gnn, l, g, g.x, e, size=(D_OUT, 4))
        assert ntrainable(gnn.GATv2Conv(2, 3)) == 5 and ntrainable(gnn.GATv2Conv((2, 4), 3, add_self_loops=False)) == 6
        assert ntrainable(gnn.GATv2Conv((2, 4), 3, add_self_loops=False, bias=False)) == 4
    else:
        assert ntrainable(gnn.GATConv(2, 3)) == 3 and ntrainable(gnn.GATConv(2, 3, bias=False)) == 2


def test_gated_graph_conv(gnn, be):                                     # conv.jl:234-245
    dev = be.dev
    l = gnn.GatedGraphConv(D_OUT, 3, aggr=gnn.mean, device=dev)
    assert tuple(l.weight.shape) == (D_OUT, D_OUT, 3)
    for g in ref_graphs(gnn, dev):
        check(gnn, l, g, g.x, size=(D_OUT, 4))


def test_agnnconv(gnn, be):                                             # conv.jl:398-415
    dev = be.dev
    l = gnn.AGNNConv(trainable=False, add_self_loops=False)
    assert l.beta.tolist() == [1.0] and l.add_self_loops is False and l.trainable is False and ntrainable(l) == 0
    l = gnn.AGNNConv(init_beta=2.0, device=dev)
    assert l.beta.tolist() == [2.0] and l.add_self_loops is True and l.trainable is True and ntrainable(l) == 1
    for g in ref_graphs(gnn, dev):
        check(gnn, l, g, g.x, size=(D_IN, 4))


@pytest.mark.parametrize("cls", ["SGConv", "TAGConv"])
def test_sg_tag_conv(gnn, be, cls):                                     # conv.jl:485-530
    dev = be.dev
    for k in (1, 2, 3):
        l = getattr(gnn, cls)(D_IN, D_OUT, k, add_self_loops=True, device=dev)
        for g in ref_graphs(gnn, dev):
            check(gnn, l, g, g.x, size=(D_OUT, 4))


def test_transformer_conv(gnn, be):                                     # conv.jl:561-591
    dev = be.dev
    ein, heads = 2, 3
    l = gnn.TransformerConv(D_IN * heads, D_IN, heads=heads, add_self_loops=True, root_weight=False, ff_channels=10,
                            skip_connection=True, batch_norm=False, device=dev)         # Kool et al., 2019
    for g in ref_graphs(gnn, dev):
        x = gnn.colmajor(torch.rand(D_IN * heads, 4, device=dev))
        check(gnn, l, g, x, size=(D_IN * heads, 4))
    l = gnn.TransformerConv((D_IN, ein), D_IN, heads=heads, gating=True, bias_qkv=True, device=dev)  # Shi et al., 2021
    for g in ref_graphs(gnn, dev):
        e = gnn.colmajor(torch.rand(ein, g.num_edges, device=dev))
        check(gnn, l, g, g.x, e, size=(D_IN * heads, 4))
    l = gnn.TransformerConv(D_IN, D_IN, heads=heads, concat=False, bias_root=False, root_weight=False, device=dev)
    for g in ref_graphs(gnn, dev):
        check(gnn, l, g, g.x, size=(D_IN, 4))
