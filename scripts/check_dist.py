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
for (n, E, D) in ((5000, 60000, 128), (200000, 3000000, 128), (30000, 200000, 64)):
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
    # partitioned
    dg = P.DistGraph(g.s, g.t, n, add_self_loops=True, device=dev)
    xl = gnn.unrows(x_full[dg.lo:dg.hi].clone()).requires_grad_(True)
    yl = P.dist_gcn_conv(layer, dg, xl)
    yl.backward(gnn.unrows(dy_full[dg.lo:dg.hi].contiguous()))
    dist.all_reduce(layer.weight.grad); dist.all_reduce(layer.bias.grad)

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp(min=1e-30))

    errs = {"y": rel(gnn.rows(yl.detach()), y_ref[dg.lo:dg.hi]), "dx": rel(gnn.rows(xl.grad), dx_ref[dg.lo:dg.hi]),
            "dW": rel(layer.weight.grad, dW_ref), "db": rel(layer.bias.grad, db_ref)}
    good = all(v < 2e-6 for v in errs.values())
    ok = ok and good
    print(f"rank {rank}/{world} n={n} E={E} D={D} range=[{dg.lo},{dg.hi}) halo_f={dg.fwd.n_halo} halo_b={dg.bwd.n_halo} "
          f"edges_f={dg.fwd.num_edges} errs={errs} {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
