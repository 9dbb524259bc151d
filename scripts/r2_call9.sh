#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "at_scale or logit or gat or lean" -p no:cacheprovider > gpurun_out/pytest_gat_lean.log 2>&1
echo "pytest gat/lean/at-scale rc=$?"; tail -n 8 gpurun_out/pytest_gat_lean.log
timeout 300 python -m pytest tests/test_layers.py tests/test_layers_reference_cases.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_layers_r2.log 2>&1
echo "pytest layers rc=$?"; tail -n 3 gpurun_out/pytest_layers_r2.log
VARIANTS=12,0,13 timeout 300 python scripts/sweep_variants.py > gpurun_out/sweep_gather4.log 2>&1
echo "sweep gather4 rc=$?"; tail -n 6 gpurun_out/sweep_gather4.log | cut -c1-300
for c in 3 4; do
    timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_c$c.log 2>&1
    echo "bench config $c rc=$?"; tail -n 1 gpurun_out/bench_c$c.log | cut -c1-1200; grep -i "Traceback" -A12 gpurun_out/bench_c$c.log | tail -n 14
done
python - <<'PY'
import json
for c in (3, 4):
    try:
        d = json.loads([l for l in open(f'gpurun_out/bench_c{c}.log').read().splitlines() if l.startswith('{')][-1])
        print(c, 'ms', d['ms_per_step'], 'kernel', d['roofline'].get('kernel_ms'), 'frac', d['roofline']['frac'], 'launch', d['roofline'].get('launch_bound'), 'parity', d['parity_rel_err'])
    except Exception as e:
        print(c, 'fail', e)
PY
