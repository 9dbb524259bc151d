"""one call of the wide tcgen05 linear kernel at config 3's projection size (5 M x 512 -> 512), for ncu"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
lib = gnn._lib.lib
N, K, Nout = 5_000_000, 512, 512
x = torch.randn(N, K, device="cuda"); W = torch.randn(Nout, K, device="cuda") / K ** 0.5
y = torch.empty(N, Nout, device="cuda")
gnn._lib.check(lib.gnnb_linear(x.data_ptr(), W.data_ptr(), None, 0, N, K, Nout, y.data_ptr(), None))
torch.cuda.synchronize()
print("tc_error", lib.gnnb_dense_tc_error())
