#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "lean" -x -p no:cacheprovider > gpurun_out/pytest_lean.log 2>&1
echo "pytest lean rc=$?"; tail -n 12 gpurun_out/pytest_lean.log
timeout 120 python scripts/debug_cheb.py > gpurun_out/debug_cheb.log 2>&1; echo "cheb rc=$?"; cat gpurun_out/debug_cheb.log | tail -n 20
VARIANTS=0,10,11,0 timeout 600 python scripts/sweep_variants.py > gpurun_out/sweep_lean.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/sweep_lean.log | tail -n 8
VARIANTS=0,10,11 timeout 600 python scripts/sweep_variants.py 5000000 50000000 256 > gpurun_out/sweep_lean_256.log 2>&1
echo "sweep256 rc=$?"; cat gpurun_out/sweep_lean_256.log | tail -n 8
GNNB_KERNEL_VARIANT=11 timeout 1000 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_v11.log 2>&1
echo "pytest v11 rc=$?"; tail -n 8 gpurun_out/pytest_gpu_v11.log
