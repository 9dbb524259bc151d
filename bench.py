#!/usr/bin/env python
"""bench.py — edges/s for forward+backward of one GCNConv 128->128 layer on a 100 M-edge RMAT graph (BASELINE.json
configs[1]), and the HBM roofline of the fused segmented-reduce kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one GCNConv forward (add_self_loops, fused 1/sqrt(d) normalised propagate, W*x, bias, relu) plus its
backward (dx, dW, db) on synthetic data: RMAT graph (seed 17), x ~ N(0,1) fp32, glorot weights.
`value`  : graph edges per second with every input resident in HBM (CUDA events, max over ranks).
`e2e`    : the same step called with HOST (pinned) arrays: x and the upstream gradient are copied host->device and
           y and dx device->host inside the timed region.
`roofline`: the dominant kernel (seg_reduce_kernel, the fused gather->message->segmented-reduce pass) timed alone
           with CUDA events; achieved = algorithmic bytes per launch / duration, against MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the oracle's restatement of the reference's CPU path (serial CSC rebuild +
           dense x CSC product, BLAS GEMM) on a bounded sample of the same workload — Julia cannot run here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# configs[1] of BASELINE.json
N_NODES, N_EDGES, DIM = 10_000_000, 100_000_000, 128
SEED = 17


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=N_NODES)
    ap.add_argument("--edges", type=int, default=N_EDGES)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--cpu-nodes", type=int, default=1_000_000, help="bounded CPU sample: nodes")
    ap.add_argument("--cpu-edges", type=int, default=10_000_000, help="bounded CPU sample: edges")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (debug)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (debug)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                               "-i", str(self.index)], text=True, timeout=5)
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_gcn_step_port(n, E, D, steps=1):
    """The reference's CPU path for one GCNConv fwd+bwd, restated (oracle = test infrastructure, timed here only as
    the baseline): add_self_loops, degree scatter, x.*c, CSC rebuild (every forward) + serial dense x CSC product,
    .*c, BLAS GEMM, bias, relu; backward = Zygote's pullbacks (Δ*A' with the forward's A, dense GEMMs)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    s, t = oracle.rmat(n, E, SEED)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, D), dtype=np.float32)
    W = (rng.standard_normal((D, D), dtype=np.float32) / np.sqrt(D)).astype(np.float32)
    b = np.zeros(D, np.float32)
    dy = rng.standard_normal((n, D), dtype=np.float32)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        s2, t2 = oracle.add_self_loops(s, t, n)                          # conv.jl:26-27
        d = oracle.degree(s2, t2, n, "in", None, np.float32)             # conv.jl:52-56
        c = (1.0 / np.sqrt(d)).astype(np.float32)
        xs = x * c[:, None]                                               # conv.jl:59
        A = oracle.csc_build(s2, t2, n, None, np.float32)                 # adjacency_matrix(g) per call, query.jl:227
        p = oracle.dense_times_csc(xs, A)                                 # xj * A, msgpass.jl:217
        p *= c[:, None]                                                   # conv.jl:67
        pre = p @ W.T + b                                                 # conv.jl:69-71
        y = np.maximum(pre, 0)
        # backward
        dpre = dy * (pre > 0)
        dW = dpre.T @ p
        db = dpre.sum(0)
        dp = dpre @ W
        dp *= c[:, None]
        dxs = oracle.dense_times_csc(dp, A, transposed=True)              # Δ * A'
        dx = dxs * c[:, None]
        times.append(time.perf_counter() - t0)
        del y, dW, db, dx
    return min(times), oracle


def cpu_leg(args):
    import numpy as np
    n, E, D = args.cpu_nodes, args.cpu_edges, args.dim
    dt, oracle = cpu_gcn_step_port(n, E, D, steps=1)
    # generous all-cores variant: prebuilt CSR + OpenMP over rows (fwd and transposed), same GEMMs
    s, t = oracle.rmat(n, E, SEED)
    s2, t2 = oracle.add_self_loops(s, t, n)
    rp, col, _ = oracle.csr(t2, s2, n)
    rpT, colT, _ = oracle.csr(s2, t2, n)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, D), dtype=np.float32)
    W = (rng.standard_normal((D, D), dtype=np.float32) / np.sqrt(D)).astype(np.float32)
    dy = rng.standard_normal((n, D), dtype=np.float32)
    c = (1.0 / np.sqrt(np.diff(rp))).astype(np.float32)
    t0 = time.perf_counter()
    p = oracle.spmm_csr_omp(rp, col, n, x, None, c, c)
    pre = p @ W.T
    dpre = dy * (pre > 0)
    dW = dpre.T @ p
    dp = dpre @ W
    dx = oracle.spmm_csr_omp(rpT, colT, n, dp, None, c, c)
    dt_omp = time.perf_counter() - t0
    del dW, dx
    return {"value": E / dt, "unit": "edges/s", "cores": 1, "kind": "port",
            "sample": f"RMAT N={n} E={E} D={D} seed {SEED}, 1 fwd+bwd GCNConv step, serial CSC-rebuild+SpMM as the "
                      f"reference (BLAS GEMM multithreaded); scaled by edges",
            "seconds": dt,
            "all_cores_openmp": {"value": E / dt_omp, "cores": oracle.num_threads(), "seconds": dt_omp,
                                 "note": "generous variant, NOT the reference's algorithm: prebuilt CSR, OpenMP rows"}}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; Julia is not installed) on host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, E, D = args.cpu_nodes, args.cpu_edges, args.dim
    t_all = []
    oracle = None
    for i in range(max(1, min(args.warmup, 1)) + max(1, min(args.steps, 3))):
        dt, oracle = cpu_gcn_step_port(n, E, D, steps=1)
        t_all.append(dt)
    timed = t_all[1:] if len(t_all) > 1 else t_all
    dt = sum(timed) / len(timed)
    val = E / dt
    line = {
        "impl": "reference", "metric": "edges/sec fwd+bwd GCNConv 128-dim (RMAT)", "value": val, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": len(t_all) - len(timed), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GCNConv {D}->{D} fwd+bwd, RMAT N={args.nodes} E={args.edges} (configs[1]); each "
                               f"reference step is a bounded sample N={n} E={E} of it"},
        "cpu_baseline": {"value": val, "unit": "edges/s", "cores": 1, "kind": "port",
                         "sample": f"RMAT N={n} E={E} D={D}, serial CSC rebuild + dense x CSC (the reference's CPU "
                                   f"algorithm; Julia unavailable so the oracle port is timed); BLAS GEMM uses "
                                   f"{os.cpu_count()} threads"},
        "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import gnnb200 as gnn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False   # fp32 GEMM like the reference (cuBLAS sgemm)
    torch.backends.cudnn.allow_tf32 = False
    n, E, D = args.nodes, args.edges, args.dim

    if world > 1:
        from gnnb200 import partition  # noqa: F401
        return run_ours_multi(args, gnn, torch, dist, world, rank, dev)

    # ---- build: graph on the device, plan, self-loop plan, transposed plans (timed separately)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = gnn.rmat_graph(n, E, SEED, device=dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.plan()
    g2 = gnn.add_self_loops(g)
    gnn._lib.check(gnn._lib.lib.gnnb_graph_csr(g2.plan().h, 1, None, None, None, None))   # transposed plan
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0

    gen = torch.Generator(device=dev).manual_seed(0)
    layer = gnn.GCNConv(D, D, torch.relu, device=dev)
    x = gnn.unrows(torch.randn(n, D, device=dev, generator=gen)).requires_grad_(True)
    dy = gnn.unrows(torch.randn(n, D, device=dev, generator=gen))

    def step():
        x.grad = None
        layer.weight.grad = None
        layer.bias.grad = None
        y = layer(g, x)
        y.backward(dy)
        return y

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = gnn.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = gnn.launch_count() - l0
    value = E / (ms * 1e-3)

    # ---- the dominant kernel alone: fused GCN propagate (forward instance), CUDA events on the launch stream
    xr = gnn.rows(x.detach())
    out = torch.empty_like(xr)
    c = gnn.layers._gcn_c(g2)
    p2 = g2.plan()
    st = torch.cuda.current_stream(dev).cuda_stream

    def kern(transposed):
        gnn._lib.check(gnn._lib.lib.gnnb_gcn_propagate(p2.h, transposed, xr.data_ptr(), None, c.data_ptr(), D,
                                                       out.data_ptr(), st))

    kt = {}
    for tr in (0, 1):
        for _ in range(3):
            kern(tr)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize()
        for a, b in evs:
            a.record(); kern(tr); b.record()
        torch.cuda.synchronize()
        kt[tr] = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    E2 = E + n
    alg_bytes = E2 * (4 * D + 4) + 4 * (n + 1) + 4 * D * n          # SURVEY.md §8d gather model, per launch
    compulsory = 4 * D * n * 2 + 4 * E2 + 4 * (n + 1)
    kms = 0.5 * (kt[0] + kt[1])
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/r1_seg_r1b_ncu_raw.csv:
    # dram__bytes_read.sum 36.83 GB + dram__bytes_write.sum 6.02 GB); only valid for the default workload
    traffic = 42.85e9 if (n, E, D) == (N_NODES, N_EDGES, DIM) else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_GBps": (traffic / (kms * 1e-3) / 1e9) if traffic else None,
            "kernel": "gnnb::seg_reduce_kernel<4,32,1,false,8,4> (fused GCN propagate, D=128)",
            "kernel_ms": {"forward": kt[0], "transposed": kt[1]}, "algorithmic_bytes_per_launch": alg_bytes,
            "compulsory_bytes_per_launch": compulsory, "peak_source": peak_src,
            "share_of_step": 2 * kms / ms}

    # ---- e2e: same step with HOST (pinned) inputs and outputs
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(n, D, pin_memory=True).normal_()
        dyh = torch.empty(n, D, pin_memory=True).normal_()
        yh = torch.empty(n, D, pin_memory=True)
        dxh = torch.empty(n, D, pin_memory=True)

        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)

        def step_host():
            # PCIe is full duplex: uploads ride stream s_in, downloads s_out, compute the main stream
            with torch.cuda.stream(s_in):
                xd_raw = xh.to(dev, non_blocking=True)
                ev_x = torch.cuda.Event(); ev_x.record(s_in)
                dyd_raw = dyh.to(dev, non_blocking=True)
                ev_dy = torch.cuda.Event(); ev_dy.record(s_in)
            main.wait_event(ev_x)
            xd = gnn.unrows(xd_raw).requires_grad_(True)
            y = layer(g, xd)
            s_out.wait_stream(main)
            with torch.cuda.stream(s_out):
                yh.copy_(gnn.rows(y.detach()), non_blocking=True)
            main.wait_event(ev_dy)
            layer.weight.grad = None
            layer.bias.grad = None
            y.backward(gnn.unrows(dyd_raw))
            s_out.wait_stream(main)
            with torch.cuda.stream(s_out):
                dxh.copy_(gnn.rows(xd.grad), non_blocking=True)
            wg = layer.weight.grad.cpu()   # D2H read of the step's result (synchronises the main stream)
            main.wait_stream(s_out)
            xd_raw.record_stream(main); dyd_raw.record_stream(main)
            return wg

        ke = max(2, min(args.steps, 5))
        step_host()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(ke):
            step_host()
        ev1.record()
        torch.cuda.synchronize()
        ems = ev0.elapsed_time(ev1) / ke
        e2e = {"value": E / (ems * 1e-3), "unit": "edges/s", "ms_per_step": ems, "steps": ke,
               "h2d_bytes_per_step": 2 * 4 * n * D, "d2h_bytes_per_step": 2 * 4 * n * D + 4 * D * D,
               "api": "gnnb200.GCNConv(g, x) on pinned host arrays: x,dy H2D; y,dx,dW D2H inside the timed region"}
        del xh, dyh, yh, dxh

    cpu = None if args.no_cpu else cpu_leg(args)
    line = {
        "metric": "edges/sec fwd+bwd GCNConv 128-dim (RMAT)", "value": value, "unit": "edges/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GCNConv {D}->{D} (add_self_loops, relu, bias) fwd+bwd on RMAT N={n} E={E} seed {SEED} "
                               f"(BASELINE configs[1]); edges counted = graph edges E (the {n} self loops are extra work)",
                   "l2": "inputs (5.1 GB features) are far larger than the 126 MB L2; no flush needed",
                   "plan_build_ms": t_plan * 1e3, "graph_gen_ms": t_gen * 1e3, "chunk_edges": 128},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def run_ours_multi(args, gnn, torch, dist, world, rank, dev):
    from gnnb200 import partition
    return partition.bench_multi(args, world, rank, dev, SEED, ClockSampler, measured_peaks)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
