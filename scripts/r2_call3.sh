#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layers_more.py -q -m gpu -k cheb -p no:cacheprovider 2>&1 | tail -n 5
VARIANTS=0 timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_lean_r2 \
    --kernel-name-base demangled -k regex:seg_lean -s 2 -c 1 python scripts/sweep_variants.py > gpurun_out/ncu_lean.log 2>&1
echo "ncu lean rc=$?"; tail -n 3 gpurun_out/ncu_lean.log
VARIANTS=10 timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_lean10_r2 \
    --kernel-name-base demangled -k regex:seg_lean -s 2 -c 1 python scripts/sweep_variants.py > gpurun_out/ncu_lean10.log 2>&1
echo "ncu lean10 rc=$?"; tail -n 3 gpurun_out/ncu_lean10.log
