#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_${N}.log 2>&1
echo "check_dist rc=$?"; grep -c " OK" gpurun_out/check_dist_${N}.log; grep "FAIL\|Error" gpurun_out/check_dist_${N}.log | head -n 5
timeout 300 $RUN --master-port 29514 scripts/halo_probe.py > gpurun_out/halo_probe_${N}.log 2>&1
echo "halo probe rc=$?"; grep "^halo" gpurun_out/halo_probe_${N}.log
