"""Times of the edge-list transforms and the neighbour sampler on the bench graph (RMAT N = 10 M, E = 100 M), CUDA events,
after one warm-up call each.  Prints one JSON line per operation with the algorithmic bytes and the achieved GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev)


def main():
    n, E = int(os.environ.get("N", 10_000_000)), int(os.environ.get("E", 100_000_000))
    g = gnn.rmat_graph(n, E, 17)
    out = {}
    ms = timed(lambda: gnn.sort_edge_index(g.s, g.t))
    out["sort_edge_index"] = {"ms": ms, "GBps": E * (16 + 16) / ms / 1e6, "bytes": "2 int64 in + 2 int64 out per edge"}
    w = torch.rand(E, device="cuda")
    gw = gnn.set_edge_weight(g, w)
    ms = timed(lambda: gnn.remove_multi_edges(gw), reps=2)
    out["remove_multi_edges(+, weights)"] = {"ms": ms}
    nodes = torch.randint(1, n + 1, (1_000_000,), device="cuda")
    g.plan()
    for K, rep in ((10, False), (10, True), (-1, False)):
        ms = timed(lambda: gnn.sample_edge_ids(g, nodes, K, replace=rep, seed=1))
        out[f"sample_edge_ids(1M nodes, K={K}, replace={rep})"] = {"ms": ms}
    ms = timed(lambda: gnn.csr(g))
    out["csr export"] = {"ms": ms}
    for k, v in out.items():
        print(json.dumps({"op": k, **v}))


if __name__ == "__main__":
    main()
