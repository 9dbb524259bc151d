"""one forward call of the tcgen05 linear kernel and one dW call at config-2 size, for ncu"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
lib = gnn._lib.lib
N, D = 10_000_000, 128
x = torch.randn(N, D, device="cuda"); W = torch.randn(D, D, device="cuda") / 11.3; b = torch.randn(D, device="cuda")
y = torch.empty(N, D, device="cuda"); dW = torch.empty(D, D, device="cuda")
for _ in range(2):
    gnn._lib.check(lib.gnnb_linear(x.data_ptr(), W.data_ptr(), b.data_ptr(), 1, N, D, D, y.data_ptr(), None))
    gnn._lib.check(lib.gnnb_linear_bwd(y.data_ptr(), None, x.data_ptr(), W.data_ptr(), 0, N, D, D, None, None, dW.data_ptr(), None, None))
torch.cuda.synchronize()
print("tc_error", lib.gnnb_dense_tc_error())
