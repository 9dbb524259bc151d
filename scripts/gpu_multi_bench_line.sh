#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${N}gpu_r2b.log 2>&1
echo "bench config 2 rc=$?"
grep "^{" gpurun_out/bench_${N}gpu_r2b.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['config']
print(d['ms_per_step'], d['value'] / 1e9, c['plan_build_ms'], c.get('plan_build_phases_ms_rank0'), c.get('nccl_connection_setup_ms'), c['per_rank']['kernel_ms'], c['per_rank']['halo_exchange_ms'], d['parity_rel_err'])"
grep -i "error\|Traceback" -A4 gpurun_out/bench_${N}gpu_r2b.log | head -n 8
