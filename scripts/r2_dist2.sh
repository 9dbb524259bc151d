#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29514 scripts/halo_probe.py > gpurun_out/halo_probe_${N}.log 2>&1
echo "halo probe rc=$?"; grep "^halo" gpurun_out/halo_probe_${N}.log
C5N=${1:-100000000}; C5E=${2:-1000000000}
timeout 900 $RUN --master-port 29513 bench.py --config 5 --gpus $N --nodes $C5N --edges $C5E --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c5_${N}gpu.log 2>&1
echo "bench config 5 (N=$C5N E=$C5E) rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_c5_${N}gpu.log').read().splitlines() if l.startswith('{')][-1])
    pr = d['config']['per_rank']
    print({'ms': d['ms_per_step'], 'Gedges_s': d['value'] / 1e9, 'plan_build_ms': d['config']['plan_build_ms'], 'kernel_ms': pr['kernel_ms'], 'halo_ms': pr['halo_exchange_ms'], 'halo_rows': pr['halo_rows_fwd'], 'halo': d['roofline']['halo'], 'frac': d['roofline']['frac']})
except Exception as e:
    print('parse failed', e)
PY
grep -i "error\|Traceback" -A3 gpurun_out/bench_c5_${N}gpu.log | head -n 12
