"""torchrun --nproc-per-node N scripts/check_dist.py : the node-partitioned GCNConv (partition.dist_gcn_conv, NCCL halo
exchange) against the single-GPU layer on the full graph: forward, dx, dW, db."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
from gnnb200 import partition as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (n, E, D) in ((5000, 60000, 128), (200000, 3000000, 128), (30000, 200000, 64), (150000, 2000000, 256)):
    g = gnn.rmat_graph(n, E, 17, device=dev)
    torch.manual_seed(0)
    layer = gnn.GCNConv(D, D, torch.relu, device=dev)
    with torch.no_grad():
        layer.bias.normal_()
    gen = torch.Generator(device=dev).manual_seed(1)
    x_full = torch.randn(n, D, device=dev, generator=gen)
    dy_full = torch.randn(n, D, device=dev, generator=gen)
    # single GPU reference (every rank computes it)
    xr = gnn.unrows(x_full.clone()).requires_grad_(True)
    y = layer(g, xr)
    y.backward(gnn.unrows(dy_full))
    y_ref, dx_ref = gnn.rows(y.detach()).clone(), gnn.rows(xr.grad).clone()
    dW_ref, db_ref = layer.weight.grad.clone(), layer.bias.grad.clone()
    layer.weight.grad = None; layer.bias.grad = None
    # partitioned: shards built on the device (csrc/shard.cu) from the resident COO and from generated chunks
    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp(min=1e-30))

    for how, ownership in (("coo", "contiguous"), ("coo", "balanced"), ("rmat-chunks", "cyclic"), ("rmat-chunks", "balanced")):
        if how == "coo":
            dg = P.DistGraph(g.s, g.t, n, add_self_loops=True, device=dev, ownership=ownership, chunk_edges=max(E // 3, 1))
        else:
            dg = P.DistGraph.from_rmat(n, E, 17, device=dev, add_self_loops=True, ownership=ownership, chunk_edges=max(E // 5, 1))
        ids = dg.local_nodes()
        xl = gnn.unrows(x_full[ids].clone()).requires_grad_(True)
        yl = P.dist_gcn_conv(layer, dg, xl)
        yl.backward(gnn.unrows(dy_full[ids].contiguous()))
        dist.all_reduce(layer.weight.grad); dist.all_reduce(layer.bias.grad)
        errs = {"y": rel(gnn.rows(yl.detach()), y_ref[ids]), "dx": rel(gnn.rows(xl.grad), dx_ref[ids]),
                "dW": rel(layer.weight.grad, dW_ref), "db": rel(layer.bias.grad, db_ref)}
        layer.weight.grad = None; layer.bias.grad = None
        good = all(v < 2e-6 for v in errs.values())
        ok = ok and good
        print(f"rank {rank}/{world} n={n} E={E} D={D} {how}/{ownership} n_local={dg.n_local} halo_f={dg.fwd.n_halo} "
              f"halo_b={dg.bwd.n_halo} edges_f={dg.fwd.num_edges} errs={errs} {'OK' if good else 'FAIL'}", flush=True)
        del dg
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
