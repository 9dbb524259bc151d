"""torchrun --nproc-per-node N scripts/halo_probe.py : the halo exchange alone (push over NVLink vs NCCL) at several row widths"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnb200 as gnn
from gnnb200 import partition as P
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, E = int(os.environ.get("N", 10_000_000)), int(os.environ.get("E", 100_000_000))
dg = P.DistGraph.from_rmat(n, E, 17, device=dev, add_self_loops=True, ownership="balanced")
for mode in ("push", "nccl"):
    os.environ["GNNB_HALO"] = mode
    for D in (64, 128, 256):
        x = torch.randn(dg.n_local, D, device=dev)
        for _ in range(3):
            dg.halo_ptr(dg.fwd, x)
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            dg.halo_ptr(dg.fwd, x)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            gb = dg.fwd.n_halo * D * 4 / 1e9
            print(f"halo {mode} D={D}: {float(t):.3f} ms for {dg.fwd.n_halo} rows = {gb:.2f} GB -> {gb / float(t) * 1e3:.0f} GB/s per GPU", flush=True)
        del x
dg.close()
dist.barrier(); dist.destroy_process_group()
