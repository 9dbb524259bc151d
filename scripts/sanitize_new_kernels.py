"""small calls of the kernels added late in round 2, meant to run under `compute-sanitizer --tool memcheck`"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
lib = gnn._lib.lib
torch.manual_seed(0)
# wide linear kernel (tail tile, 3 output quarters, 5 K-blocks) + its dx
N, K, Nout = 2048 + 77, 160, 384
x = torch.randn(N, K, device="cuda"); W = torch.randn(Nout, K, device="cuda") / K ** 0.5; b = torch.randn(Nout, device="cuda")
y = torch.empty(N, Nout, device="cuda")
gnn._lib.check(lib.gnnb_linear(x.data_ptr(), W.data_ptr(), b.data_ptr(), 1, N, K, Nout, y.data_ptr(), None))
ref = (x.double() @ W.double().t() + b.double()).clamp(min=0)
print("wide linear err", float((y.double() - ref).norm() / ref.norm()), "tc_error", lib.gnnb_dense_tc_error())
# closing line
D = 36
z = torch.randn(1001, D, device="cuda"); bb = torch.randn(D, device="cuda"); o = torch.empty_like(z)
gnn._lib.check(lib.gnnb_bias_act(z.data_ptr(), bb.data_ptr(), 1, 1001, D, o.data_ptr(), None))
dz = torch.empty_like(z); db = torch.empty(D, device="cuda")
gnn._lib.check(lib.gnnb_bias_act_bwd(z.data_ptr(), o.data_ptr(), 1, 1001, D, dz.data_ptr(), db.data_ptr(), None))
print("bias_act ok", bool(torch.equal(o, (z + bb).clamp(min=0))))
# max pullback on a graph with a hub in the by-source plan (long rows -> partial slots + fix-up)
n = 3000
s = torch.cat([torch.ones(1500, dtype=torch.int64), torch.randint(1, n + 1, (4000,))])
t = torch.cat([torch.randint(1, n + 1, (1500,)), torch.randint(1, n + 1, (4000,))])
g = gnn.GNNGraph(s, t, num_nodes=n).to("cuda")
for Dm in (128, 256):
    xm = gnn.unrows((torch.randn(n, Dm, device="cuda") * 2).round() / 2).requires_grad_(True)
    ym = gnn.propagate(gnn.copy_xj, g, "max", xj=xm)
    dy = torch.where(torch.isfinite(ym), torch.randn_like(ym), torch.zeros_like(ym))
    ym.backward(dy)
    print("max pullback D", Dm, float(xm.grad.abs().sum()))
torch.cuda.synchronize()
print("done")
