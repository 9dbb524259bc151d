#!/bin/bash
# round-2 first GPU call: whole GPU suite with the gate removed, transform/sampler timings, prefetch-variant sweep
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 1000 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_r2.log 2>&1
echo "pytest -m gpu rc=$?"; tail -n 15 gpurun_out/pytest_gpu_r2.log
timeout 300 python scripts/time_transforms.py > gpurun_out/time_transforms.log 2>&1
echo "time_transforms rc=$?"; cat gpurun_out/time_transforms.log | tail -n 12
VARIANTS=0,6,7,8,9,0 timeout 600 python scripts/sweep_variants.py > gpurun_out/sweep_prefetch.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/sweep_prefetch.log | tail -n 12
