#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "at_scale or logit or gat" -x -p no:cacheprovider > gpurun_out/pytest_scale.log 2>&1
echo "pytest at-scale/gat rc=$?"; tail -n 8 gpurun_out/pytest_scale.log
timeout 300 python -m pytest tests/test_layers.py tests/test_layers_reference_cases.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_layers_r2.log 2>&1
echo "pytest layers rc=$?"; tail -n 4 gpurun_out/pytest_layers_r2.log
for c in 1 4 3; do
    timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_c$c.log 2>&1
    echo "bench config $c rc=$?"; tail -n 1 gpurun_out/bench_c$c.log | cut -c1-2500; grep -i "Traceback" -A12 gpurun_out/bench_c$c.log | tail -n 14
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2>&1
echo "bench config 2 rc=$?"; tail -n 1 gpurun_out/bench_c2.log | cut -c1-3000; grep -i "Traceback" -A12 gpurun_out/bench_c2.log | tail -n 14
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py -q -m gpu -k "linear or gcn or reparam" -x -p no:cacheprovider > gpurun_out/pytest_linear.log 2>&1
echo "pytest linear/gcn rc=$?"; tail -n 4 gpurun_out/pytest_linear.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 launch list rc=$?"; tail -n 2 gpurun_out/ncu_c3.log | cut -c1-300
