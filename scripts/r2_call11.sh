#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_layers.py -q -m gpu -k "linear2 or sage or split or lean" -p no:cacheprovider > gpurun_out/pytest_linear2.log 2>&1
echo "pytest linear2/sage rc=$?"; tail -n 5 gpurun_out/pytest_linear2.log
VARIANTS=12,0,13 timeout 300 python scripts/sweep_variants.py > gpurun_out/sweep_gather4_128.log 2>&1; tail -n 3 gpurun_out/sweep_gather4_128.log | cut -c1-200
VARIANTS=12,0,13 timeout 300 python scripts/sweep_variants.py 5000000 50000000 256 > gpurun_out/sweep_gather4_256.log 2>&1; tail -n 3 gpurun_out/sweep_gather4_256.log | cut -c1-200
VARIANTS=12,0,13 timeout 300 python scripts/sweep_variants.py 2500000 25000000 512 > gpurun_out/sweep_gather4_512.log 2>&1; tail -n 3 gpurun_out/sweep_gather4_512.log | cut -c1-200
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/bench_c4.log 2>&1
echo "bench config 4 rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_c4.log').read().splitlines() if l.startswith('{')][-1])
print('c4 ms', d['ms_per_step'], d['roofline'].get('launch_bound'), d['parity_rel_err'])
PY
VARIANTS=13 timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_gather4_r2 --kernel-name-base demangled -k regex:seg_gather4 -s 2 -c 1 python scripts/sweep_variants.py > gpurun_out/ncu_gather4.log 2>&1
echo "ncu gather4 rc=$?"; tail -n 2 gpurun_out/ncu_gather4.log | cut -c1-200
