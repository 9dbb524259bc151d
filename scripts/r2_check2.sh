#!/bin/bash
set -u
mkdir -p gpurun_out /tmp/rep
GNNB_BENCH_PARTITIONED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29520 bench.py --config 2 --nodes 2000000 --edges 20000000 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_partitioned_1rank.log 2>&1
echo "partitioned path on one rank rc=$?"; tail -n 1 gpurun_out/bench_partitioned_1rank.log | python -c "import sys, json; l = sys.stdin.read(); print(json.loads(l)['parity_rel_err'] if l.startswith('{') else l[-1500:])"
timeout 200 python scripts/time_maxmin_bwd.py > gpurun_out/time_maxmin_bwd.log 2>&1; echo "maxmin rc=$?"; cat gpurun_out/time_maxmin_bwd.log | tail -n 4
timeout 300 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/prof_lean_mean_c4_r2 --kernel-name-base demangled -k regex:'seg_lean_kernel<\(int\)1, \(int\)0, \(bool\)0, \(int\)0, \(int\)1>' -s 2 -c 1 python bench.py --config 4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
[ -f /tmp/rep/prof_lean_mean_c4_r2.ncu-rep ] && python scripts/ncu_raw_extract.py /tmp/rep/prof_lean_mean_c4_r2.ncu-rep > gpurun_out/prof_lean_mean_c4_r2_ncu_raw.csv
ls -la /tmp/rep
timeout 300 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/prof_wide_r2 --kernel-name-base demangled -k regex:'linear_wide' -c 1 python scripts/wide_linear_one.py > gpurun_out/ncu_wide.log 2>&1
echo "ncu wide rc=$?"
[ -f /tmp/rep/prof_wide_r2.ncu-rep ] && python scripts/ncu_raw_extract.py /tmp/rep/prof_wide_r2.ncu-rep > gpurun_out/prof_wide_r2_ncu_raw.csv && cp /tmp/rep/prof_wide_r2.ncu-rep gpurun_out/
timeout 400 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/bench_c4_final.log 2>&1; echo "bench c4 rc=$?"; tail -n 1 gpurun_out/bench_c4_final.log | cut -c1-300
