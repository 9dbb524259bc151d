"""One launch (after one warm-up) of every kernel family of libgnnb200 at a size larger than L2, for `ncu`:

    ncu --set full --clock-control none --import-source on -f -o gpurun_out/all_kernels \
        --kernel-name-base demangled -k regex:'gnnb|tc::|tcw::' python scripts/run_all_kernels.py
    python scripts/ncu_summarize.py gpurun_out/all_kernels.ncu-rep > profiles/r2_all_kernels.md

Sizes: RMAT N = 2 M, E = 20 M (+ self loops), D = 128 (1 GB of features; a fifth of the bench graph so that the ~40
replays per kernel of `--set full` stay within a few minutes).  NVTX-free: the kernel names identify the family."""
import operator
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn  # noqa: E402


def main():
    n, E, D = int(os.environ.get("N", 2_000_000)), int(os.environ.get("E", 20_000_000)), 128
    dev = os.environ.get("DEV", "cuda")                            # DEV=cpu: dry run on tests/fake_abi.py (no kernels)
    if dev == "cuda":
        g = gnn.rmat_graph(n, E, 17)                               # rmat_kernel
    else:
        g = gnn.GNNGraph(torch.randint(1, n + 1, (E,)), torch.randint(1, n + 1, (E,)), num_nodes=n)
    x = gnn.unrows(torch.randn(n, D, device=dev)).requires_grad_(True)
    w = torch.rand(E, device=dev)
    for rep in range(2):                                           # warm-up, then the captured pass (use `-s` to skip)
        gl = gnn.add_self_loops(g)                                 # plan build + self-loop derivation
        for aggr in (operator.add, gnn.mean, max):
            y = gnn.propagate(gnn.copy_xj, g, aggr, xj=x)          # seg_reduce (+ fix-up), three reductions
            y.sum().backward()                                     # transposed pass / maxmin_bwd
        y = gnn.propagate(gnn.e_mul_xj, g, operator.add, xj=x, e=w.requires_grad_(True))
        y.sum().backward()                                         # weighted + edge_dot (dw)
        layer = gnn.GCNConv(D, D, torch.relu, device=dev)
        layer(g, x).sum().backward()                               # gcn_norm, gcn propagate x2, tcgen05 linear / dx / dW, act_bwd
        m = gnn.apply_edges(gnn.xi_sub_xj, g, xi=x, xj=x)          # gather x2
        gnn.aggregate_neighbors(g, max, m).sum().backward()        # scatter (edge-id view) + its pullback
        e = gnn.unrows(torch.randn(E, 8, device=dev)).requires_grad_(True)
        gnn.softmax_edge_neighbors(g, e).sum().backward()          # edge softmax fwd / bwd
        gat = gnn.GATConv(D, 64, heads=8, device=dev)
        gat(g, x).sum().backward()                                 # gat_fwd / gat_bwd (+ fix-ups)
        gnn.sort_edge_index(g.s, g.t)                              # encode / radix sort / decode
        gnn.remove_multi_edges(gnn.set_edge_weight(g, w.detach())) # coalesce + segmented scatter
        nodes = torch.randint(1, n + 1, (200_000,), device=dev)
        gnn.sample_edge_ids(g, nodes, 10, seed=rep)                # sampler
        gnn.csr(g, True)
        if dev == "cuda":
            torch.cuda.synchronize()
    print("launches", gnn.launch_count())


if __name__ == "__main__":
    if os.environ.get("DEV") == "cpu":
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import fake_abi
        with fake_abi.installed():
            main()
    else:
        main()
