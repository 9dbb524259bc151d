#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "lean or hot or halo or tma" -x -p no:cacheprovider > gpurun_out/pytest_hot.log 2>&1
echo "pytest hot rc=$?"; tail -n 6 gpurun_out/pytest_hot.log
VARIANTS=0 HOT=0,32,64,96 MODES=1,2,3 timeout 900 python scripts/sweep_variants.py > gpurun_out/sweep_hot.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/sweep_hot.log | tail -n 14
