"""The checker behind the multi-GPU bench lines' `parity_rel_err` (bench.py: dist_parity), itself checked on the CPU.

dist_parity runs unchanged over a stand-in partitioned graph whose propagate is a dense float64 product on a PERMUTED node
order (so the local-row <-> global-id mapping matters) and whose edge list comes in two chunks.  With a correct stand-in it
must report errors at fp32 round-off; with a stand-in that forgets the self loops, or maps rows to the wrong nodes, it must
report a large error.  One gloo rank; the oracle restates the sampled rows exactly as on the GPU box."""
import importlib.util
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_parity_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _StandIn:
    """what dist_parity needs of a DistGraph, on one CPU rank"""

    def __init__(self, s1, t1, n, perm, self_loops=True, shift_rows=False):
        self.device, self.rank, self.world, self.n_local = torch.device("cpu"), 0, 1, n
        self.perm = perm                                          # local row i holds global node perm[i]
        self.fwd, self.bwd = "fwd", "bwd"
        s, t = s1 - 1, t1 - 1
        A = torch.zeros(n, n, dtype=torch.float64)
        A.index_put_((t, s), torch.ones(s.numel(), dtype=torch.float64), accumulate=True)   # A[target, source]
        deg = torch.bincount(t, minlength=n).double()
        if self_loops:
            A += torch.eye(n, dtype=torch.float64)
        self.c = 1.0 / torch.sqrt(deg + 1.0)                                                # the layer's normalisation
        self.A = A
        self.shift = shift_rows

    def local_nodes(self):
        return self.perm.roll(1) if self.shift else self.perm     # a wrong mapping when shift_rows

    def gcn_c(self):
        c = self.c[self.perm].float()
        return c, c, c

    def propagate(self, shard, x_rows, cs, ct):
        M = self.A if shard == "fwd" else self.A.t()
        xg = torch.zeros(self.n_local, x_rows.shape[1], dtype=torch.float64)
        xg[self.perm] = x_rows.double()                           # local order -> global order
        out = self.c[:, None] * (M @ (self.c[:, None] * xg))
        return out[self.perm].float()


@pytest.fixture
def one_gloo_rank():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("fault", [None, "no_self_loops", "wrong_row_mapping"])
def test_dist_parity_checker(one_gloo_rank, monkeypatch, fault):
    bench = _bench()
    import gnnb200 as gnn
    from gnnb200 import partition
    torch.manual_seed(3)
    n, E, D = 500, 4000, 8
    s1 = torch.randint(1, n + 1, (E,), dtype=torch.int64)
    t1 = torch.randint(1, n + 1, (E,), dtype=torch.int64)
    s1[:7] = t1[:7]                                               # a few self loops already in the list
    dg = _StandIn(s1, t1, n, torch.randperm(n), self_loops=fault != "no_self_loops", shift_rows=fault == "wrong_row_mapping")
    layer = SimpleNamespace(weight=torch.randn(D, D) / D ** 0.5, bias=torch.randn(D))

    def chunks(num_nodes, num_edges, seed, device, chunk_edges):
        assert (num_nodes, num_edges) == (n, E)
        yield s1[:E // 3], t1[:E // 3]
        yield s1[E // 3:], t1[E // 3:]

    def layer_forward(l, g, x):                                   # sigma.(W * propagate(x) .+ b) on Julia-shaped x
        c, cf, cb = g.gcn_c()
        pr = g.propagate(g.fwd, gnn.rows(x), cf, c)
        return gnn.unrows(torch.relu(pr @ l.weight.t() + l.bias))

    monkeypatch.setattr(partition, "rmat_chunks", chunks)
    monkeypatch.setattr(partition, "dist_gcn_conv", layer_forward)
    res = bench.dist_parity(SimpleNamespace(nodes=n, edges=E, dim=D), dg, layer, samples=60, max_deg=10_000)
    assert res["rows_checked"] > 50 and res["rows_dropped_for_degree"] == 0 and res["edges_restated"] > 500
    assert res["adjoint_identity_all_rows"] < 1e-6               # holds for any symmetric pair of shards, faulty or not
    worst = max(res["propagate_rows"], res["layer_forward_rows"], res["transposed_propagate_rows"])
    if fault is None:
        assert worst < 5e-6
    else:
        assert res["propagate_rows"] > 1e-2 and res["transposed_propagate_rows"] > 1e-2
