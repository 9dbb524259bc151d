#!/bin/bash
set -u
mkdir -p gpurun_out /tmp/rep
timeout 200 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/prof_wide_v2_r2 --kernel-name-base demangled -k regex:'linear_wide' -c 1 python scripts/wide_linear_one.py > gpurun_out/ncu_wide_v2.log 2>&1
echo "ncu wide rc=$?"
[ -f /tmp/rep/prof_wide_v2_r2.ncu-rep ] && python scripts/ncu_raw_extract.py /tmp/rep/prof_wide_v2_r2.ncu-rep > gpurun_out/prof_wide_v2_r2_ncu_raw.csv
timeout 120 python bench.py --config 1 --steps 20 --warmup 5 > gpurun_out/bench_c1_final.log 2>&1; echo "bench c1 rc=$?"; tail -n 1 gpurun_out/bench_c1_final.log | cut -c1-250
