#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, cheapest first.
#   gpurun --timeout 900 -- 'bash scripts/r2_first_gpu_run.sh'            (1 GPU part)
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/r2_first_gpu_run.sh dist'   (overlapped halo schedule, A/B)
set -u
mkdir -p gpurun_out
if [ "${1:-}" = "dist" ]; then
    N=$(nvidia-smi -L | wc -l)
    for ov in 0 1; do
        GNNB_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
            --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_ov$ov.log 2>&1
        echo "check_dist overlap=$ov rc=$?"; tail -n 3 gpurun_out/check_dist_ov$ov.log
        GNNB_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
            --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_${N}gpu_ov$ov.log 2>&1
        echo "bench overlap=$ov rc=$?"; tail -n 1 gpurun_out/bench_${N}gpu_ov$ov.log | cut -c1-400
    done
    exit 0
fi
# 1. the gated CUDA cases of the neighbour sampler (csrc/sample.cu has never run on a GPU)
GNNB_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_sampling.py -q -m gpu -p no:cacheprovider > gpurun_out/sampling_cuda.log 2>&1
echo "sampling rc=$?"; tail -n 5 gpurun_out/sampling_cuda.log
# 2. the whole GPU suite on the rebuilt library
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu_r2.log 2>&1
echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2.log
# 3. time the transforms and the sampler at config-2 size
timeout 300 python scripts/time_transforms.py > gpurun_out/time_transforms.log 2>&1
echo "time_transforms rc=$?"; cat gpurun_out/time_transforms.log
