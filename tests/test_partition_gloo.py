"""Host-side logic of the node-partitioned path (partition.py) on CPU with the gloo backend, world_size 2 and 3:
cost-balanced bounds, shard re-indexing into [local | halo], deduplicated request lists, the all-to-all-v halo
exchange — checked by emulating the fused kernel with torch index_add and comparing the assembled result with the
full-graph product (forward shard) and its transpose (backward shard).  No GPU, no libgnnb200 compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _emulate(shard, x_local, halo_rows, cs, ct):
    src = torch.cat([x_local, halo_rows]) * cs[:, None]
    out = torch.zeros_like(x_local)
    out.index_add_(0, shard._d["row"], src[shard._d["col"]])
    return out * ct[:, None]


def _worker(rank, world, port, n, E, seed, q, ownership="contiguous"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gnnb200
    from gnnb200 import partition as P
    rng = np.random.default_rng(seed)
    # skewed graph: low ids are hubs (like RMAT), duplicates and self loops kept
    s = np.minimum((rng.random(E) ** 3 * n).astype(np.int64), n - 1) + 1
    t = np.minimum((rng.random(E) ** 2 * n).astype(np.int64), n - 1) + 1
    D = 5
    x = torch.as_tensor(rng.standard_normal((n, D)))
    # keep the raw shard dicts for the emulation
    orig = P.build_shard
    kept = []

    def spy(*a, **k):
        d = orig(*a, **k)
        kept.append(d)
        return d

    P.build_shard = spy
    dg = P.DistGraph(torch.as_tensor(s), torch.as_tensor(t), n, add_self_loops=True, device="cpu", ownership=ownership)
    dg.fwd._d, dg.bwd._d = kept[0], kept[1]
    b = dg.bounds
    assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
    lo, hi = dg.lo, dg.hi
    ids = dg.local_nodes()                                   # node of every local row
    node_of = torch.argsort(P.to_pid(torch.arange(n), world, dg.first, ownership, dg._relabel))   # partition id -> node
    assert torch.equal(node_of[lo:hi], ids)
    # dedup: each remote row requested once; grouped by owner in rank order
    for sh in (dg.fwd, dg.bwd):
        halo = sh._d["halo"]
        assert (halo[1:] > halo[:-1]).all() and ((halo < lo) | (halo >= hi)).all()
        assert sum(sh.recv_counts) == halo.numel() and sh.recv_counts[rank] == 0
        assert sh.send_idx.numel() == sum(sh.send_counts) and ((sh.send_idx >= 0) & (sh.send_idx < hi - lo)).all()
    # degrees incl. self loops -> c, exchanged to the halos
    s0, t0 = torch.as_tensor(s) - 1, torch.as_tensor(t) - 1
    deg = torch.bincount(t0, minlength=n).double() + 1
    c_full = 1 / deg.sqrt()
    c = c_full[ids]
    cf = torch.cat([c, dg.halo(dg.fwd, c.reshape(-1, 1)).reshape(-1)])
    cb = torch.cat([c, dg.halo(dg.bwd, c.reshape(-1, 1)).reshape(-1)])
    assert torch.equal(cf[dg.n_local:], c_full[node_of[dg.fwd._d["halo"]]])
    xl = x[ids].contiguous()
    out_f = _emulate(dg.fwd, xl, dg.halo(dg.fwd, xl), cf, c)
    out_b = _emulate(dg.bwd, xl, dg.halo(dg.bwd, xl), cb, c)
    # full-graph reference: A_hat = C (A + I) C
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((s0, t0), torch.ones(E, dtype=torch.float64), accumulate=True)
    A += torch.eye(n, dtype=torch.float64)
    ref_f = (c_full[:, None] * (A.t() @ (c_full[:, None] * x)))[ids]
    ref_b = (c_full[:, None] * (A @ (c_full[:, None] * x)))[ids]
    ok = bool(torch.allclose(out_f, ref_f, rtol=1e-12, atol=1e-12) and torch.allclose(out_b, ref_b, rtol=1e-12, atol=1e-12))
    costs = torch.tensor([float((torch.bincount(t0, minlength=n) + torch.bincount(s0, minlength=n) + P.NODE_COST)[ids].sum())])
    allc = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(allc, costs)
    q.put((rank, ok, [float(v) for v in allc], dg.fwd.num_edges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ownership", [(2, "contiguous"), (3, "contiguous"), (2, "cyclic"), (3, "cyclic"),
                                             (2, "balanced"), (3, "balanced")])
def test_partition_halo_logic_gloo(world, ownership):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, E = 61, 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, E, 3, q, ownership)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    costs = res[0][2]
    if ownership != "cyclic":
        assert max(costs) < (1.6 if ownership == "contiguous" else 1.25) * (sum(costs) / world)   # balanced cost per rank
    assert sum(e for _, _, _, e in res) == E + n               # every edge (and self loop) is owned exactly once


def test_balanced_bounds_edge_cases():
    import gnnb200
    from gnnb200.partition import balanced_bounds
    assert balanced_bounds(torch.ones(10), 1) == [0, 10]
    b = balanced_bounds(torch.ones(10), 2)
    assert b == [0, 5, 10]
    b = balanced_bounds(torch.tensor([100.0, 1, 1, 1]), 2)       # one hub: it gets a range of its own
    assert b[0] == 0 and b[-1] == 4 and b[1] in (1, 2)
    b = balanced_bounds(torch.ones(3), 8)                         # more ranks than nodes: empty ranges allowed
    assert b[0] == 0 and b[-1] == 3 and all(b[i] <= b[i + 1] for i in range(8))


def _worker_layer(rank, world, port, n, E, seed, q, slices="1"):
    """dist_gcn_conv forward + backward under gloo, the library replaced by the CPU test double (tests/fake_abi.py):
    the whole host path of the multi-GPU layer — shard plans, c = 1/sqrt(d) with its halo copies, the two halo
    exchanges, the autograd function, the gradient all-reduce — against the dense full-graph formula."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GNNB_HALO="nccl", GNNB_HALO_SLICES=slices)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import fake_abi
    import gnnb200 as gnn
    from gnnb200 import partition as P
    rng = np.random.default_rng(seed)
    s = np.minimum((rng.random(E) ** 3 * n).astype(np.int64), n - 1) + 1
    t = np.minimum((rng.random(E) ** 2 * n).astype(np.int64), n - 1) + 1
    Din, Dout = 6, 4
    x_full = rng.standard_normal((n, Din)).astype(np.float32)
    dy_full = rng.standard_normal((n, Dout)).astype(np.float32)
    ok, err = False, ""
    with fake_abi.installed() as fake:
        kept, orig = [], P.build_shard
        P.build_shard = lambda *a, **k: (kept.append(orig(*a, **k)) or kept[-1])
        dg = P.DistGraph(torch.as_tensor(s), torch.as_tensor(t), n, add_self_loops=True, device="cpu")
        P.build_shard = orig
        for sh, d in ((dg.fwd, kept[0]), (dg.bwd, kept[1])):       # the plans csrc/shard.cu creates on a CUDA device
            sh.plan = dg._plans(d)
        torch.manual_seed(0)
        layer = gnn.GCNConv(Din, Dout, torch.relu)
        with torch.no_grad():
            layer.bias.copy_(torch.linspace(-0.5, 0.5, Dout))
        lo, hi = dg.lo, dg.hi
        x = gnn.unrows(torch.as_tensor(x_full[lo:hi]).contiguous()).requires_grad_(True)
        y = P.dist_gcn_conv(layer, dg, x)
        y.backward(gnn.unrows(torch.as_tensor(dy_full[lo:hi]).contiguous()))
        dist.all_reduce(layer.weight.grad)
        dist.all_reduce(layer.bias.grad)
        # dense reference on the full graph, float64
        s0, t0 = torch.as_tensor(s) - 1, torch.as_tensor(t) - 1
        A = torch.zeros(n, n, dtype=torch.float64)
        A.index_put_((s0, t0), torch.ones(E, dtype=torch.float64), accumulate=True)
        A += torch.eye(n, dtype=torch.float64)
        c = 1 / A.sum(0).sqrt()
        xr = torch.as_tensor(x_full, dtype=torch.float64).requires_grad_(True)
        W = layer.weight.detach().double().requires_grad_(True)
        b = layer.bias.detach().double().requires_grad_(True)
        yr = torch.relu((c[:, None] * (A.t() @ (c[:, None] * xr))) @ W.t() + b)
        yr.backward(torch.as_tensor(dy_full, dtype=torch.float64))
        close = lambda a, r: bool(torch.allclose(a.double(), r, rtol=2e-5, atol=2e-6))
        checks = {"y": close(gnn.rows(y), yr[lo:hi]), "dx": close(gnn.rows(x.grad), xr.grad[lo:hi]),
                  "dW": close(layer.weight.grad, W.grad), "db": close(layer.bias.grad, b.grad),
                  "calls": (fake.calls.count("gnnb_propagate_halo") == 2 * int(slices)
                            and fake.calls.count("gnnb_gcn_norm") == 1)}
        ok = all(checks.values())
        err = str(checks)
    q.put((rank, ok, err))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,slices", [(2, "1"), (3, "1"), (2, "2")])
def test_dist_gcn_conv_gloo_on_the_test_double(world, slices):
    """slices = "2": the column-sliced exchange (GNNB_HALO_SLICES)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_layer, args=(r, world, port, 50, 400, 5, q, slices))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
