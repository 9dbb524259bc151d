#!/bin/bash
# the 8-GPU call: config 5 (1 B edges, D = 256) first, then the bench line (config 2), the parity check of the partitioned
# layer against the single-GPU layer, the halo probe.  Every step has its own timeout; the call as a whole is bounded by gpurun.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    pr = d['config']['per_rank']
    print({'ms': round(d['ms_per_step'], 3), 'Gedges_s': round(d['value'] / 1e9, 3), 'plan_build_ms': round(d['config']['plan_build_ms']),
           'kernel_ms': [round(v, 2) for v in pr['kernel_ms']], 'halo_ms': [round(v, 2) for v in pr['halo_exchange_ms']],
           'halo_rows': pr['halo_rows_fwd'], 'shard_edges': pr['shard_edges'], 'halo_GBps': round(d['roofline']['halo']['GBps_per_gpu']),
           'frac': round(d['roofline']['frac'], 3), 'e2e': d.get('e2e') and round(d['e2e']['ms_per_step'], 1)})
    print('parity', d.get('parity_rel_err'))
except Exception as e:
    print('parse failed', e)
PY
}
C5N=${1:-100000000}; C5E=${2:-1000000000}
timeout 300 $RUN --master-port 29513 bench.py --config 5 --gpus $N --nodes $C5N --edges $C5E --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c5_${N}gpu.log 2>&1
echo "bench config 5 (N=$C5N E=$C5E) rc=$?"; summ gpurun_out/bench_c5_${N}gpu.log; grep -i "error\|Traceback" -A4 gpurun_out/bench_c5_${N}gpu.log | head -n 16
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -n 2
timeout 150 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu_r2.log 2>&1
echo "bench config 2 rc=$?"; summ gpurun_out/bench_${N}gpu_r2.log; grep -i "error\|Traceback" -A4 gpurun_out/bench_${N}gpu_r2.log | head -n 8
timeout 150 $RUN --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_${N}.log 2>&1
echo "check_dist rc=$?"; grep -c " OK" gpurun_out/check_dist_${N}.log; grep "FAIL\|Error" gpurun_out/check_dist_${N}.log | head -n 5
timeout 80 $RUN --master-port 29514 scripts/halo_probe.py > gpurun_out/halo_probe_${N}.log 2>&1
echo "halo probe rc=$?"; grep "^halo" gpurun_out/halo_probe_${N}.log
