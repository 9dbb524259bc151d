import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
lib = gnn._lib.lib
for (N, Din, Dout) in ((37, 128, 128), (4096, 64, 128), (100000, 32, 128), (1000000, 128, 128), (10000000, 128, 128)):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, Din, device="cuda", generator=g)
    W = torch.randn(Dout, Din, device="cuda", generator=g) / Din ** 0.5
    dy = torch.randn(N, Dout, device="cuda", generator=g)
    dW = torch.full((Dout, Din), 7.0, device="cuda")
    rc = lib.gnnb_linear_bwd(dy.data_ptr(), None, x.data_ptr(), W.data_ptr(), 0, N, Din, Dout, None, None, dW.data_ptr(), None, None)
    torch.cuda.synchronize()
    ref = (dy.double().t() @ x.double())
    err = float((dW.double() - ref).norm() / ref.norm())
    print(N, Din, Dout, "rc", rc, "tc_error", lib.gnnb_dense_tc_error(), "rel err", err, "dW[0,:4]", dW[0, :4].tolist(), "ref", ref[0, :4].tolist(), flush=True)
