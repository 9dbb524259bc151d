"""forward + pullback of max aggregation at N = 2 M / E = 20 M (RMAT), D = 128: CUDA events"""
import operator, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
n, E, D = 2_000_000, 20_000_000, int(os.environ.get("D", "128"))
g = gnn.rmat_graph(n, E, 17)
x = gnn.unrows(torch.randn(n, D, device="cuda")).requires_grad_(True)
dy = gnn.unrows(torch.randn(n, D, device="cuda"))
for aggr in ("max", "mean"):
    for _ in range(2):
        y = gnn.propagate(gnn.copy_xj, g, aggr, xj=x)
        dy2 = torch.where(torch.isfinite(y), dy, torch.zeros_like(dy))
        x.grad = None
        y.backward(dy2)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    y = gnn.propagate(gnn.copy_xj, g, aggr, xj=x)
    e1.record()
    x.grad = None
    y.backward(dy2)
    e2.record()
    torch.cuda.synchronize()
    print(f"{aggr} D={D}: forward {e0.elapsed_time(e1):.3f} ms, pullback {e1.elapsed_time(e2):.3f} ms", flush=True)
