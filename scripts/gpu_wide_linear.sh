#!/bin/bash
# the wide tcgen05 linear kernel and the fused closing line: correctness first, then time, then config 3 with both
set -u
mkdir -p gpurun_out
GNNB_BENCH_PARTITIONED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29520 bench.py --config 2 --nodes 2000000 --edges 20000000 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_partitioned_1rank.log 2>&1
echo "partitioned path on one rank rc=$?"; tail -n 1 gpurun_out/bench_partitioned_1rank.log | python -c "import sys, json; l = sys.stdin.read(); print(json.loads(l)['parity_rel_err'] if l.startswith('{') else l[-1500:])"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "linear or bias_act or closing_line or gradients" > gpurun_out/pytest_wide.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_wide.log
timeout 300 python scripts/wide_linear_check.py > gpurun_out/wide_linear_check.log 2>&1
echo "wide check rc=$?"; cat gpurun_out/wide_linear_check.log | tail -n 8
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/bench_c3_wide.log 2>&1
echo "bench config 3 rc=$?"; tail -n 1 gpurun_out/bench_c3_wide.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_wide.csv python bench.py --config 3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c3_wide.log 2>&1
echo "ncu config 3 launch list rc=$?"
timeout 600 python -m pytest tests -q -x -p no:cacheprovider -m gpu -k "gat or layers or scale or sage" > gpurun_out/pytest_wide2.log 2>&1
echo "pytest gat/layers rc=$?"; tail -n 3 gpurun_out/pytest_wide2.log
