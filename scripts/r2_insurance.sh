#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke_last.log
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_last.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_last.log | cut -c1-200
