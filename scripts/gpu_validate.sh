#!/bin/bash
# final single-GPU validation: whole GPU suite, smoke, the driver's bench command, the wide kernel's check, config 3
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest -m gpu rc=$?"; tail -n 4 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_1gpu.log 2>&1
echo "bench rc=$?"; tail -n 1 gpurun_out/bench_final_1gpu.log | cut -c1-300
timeout 200 python scripts/wide_linear_check.py > gpurun_out/wide_linear_check2.log 2>&1
echo "wide check rc=$?"; tail -n 6 gpurun_out/wide_linear_check2.log
timeout 400 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/bench_c3_final.log 2>&1
echo "bench config 3 rc=$?"; tail -n 1 gpurun_out/bench_c3_final.log | cut -c1-300
GNNB_BENCH_PARTITIONED=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29520 bench.py --config 2 --nodes 2000000 --edges 20000000 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_partitioned_1rank.log 2>&1
echo "partitioned path on one rank rc=$?"; tail -n 1 gpurun_out/bench_partitioned_1rank.log | python -c "import sys, json; l = sys.stdin.read(); d = json.loads(l); print(d['parity_rel_err']['layer_forward_rows'], d['config']['plan_build_ms'], d['config'].get('plan_build_phases_ms_rank0'), d['config'].get('nccl_connection_setup_ms'))"
