#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, cheapest first.
#   gpurun --timeout 900 -- 'bash scripts/r2_first_gpu_run.sh'            (1 GPU part)
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/r2_first_gpu_run.sh dist'   (overlapped halo schedule, A/B)
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/r2_first_gpu_run.sh c5'     (BASELINE configs[4]: 1 B edges, D = 256)
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
if [ "${1:-}" = "c5" ]; then
    # never run in round 1: every rank generates the 1 B-edge RMAT list (16 GB) and keeps its shard
    GNNB_HALO_SLICES=4 timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        bench.py --gpus 8 --nodes 100000000 --edges 1000000000 --dim 256 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c5.log 2>&1
    echo "c5 rc=$?"; tail -n 2 gpurun_out/bench_c5.log | cut -c1-600
    exit 0
fi
# 1. the gated CUDA cases of the neighbour sampler (csrc/sample.cu has never run on a GPU)
GNNB_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_sampling.py tests/test_layers_reference_cases.py tests/test_layers_more.py tests/test_query.py -q -m gpu -p no:cacheprovider > gpurun_out/sampling_cuda.log 2>&1
echo "sampling rc=$?"; tail -n 5 gpurun_out/sampling_cuda.log
# 2. the whole GPU suite on the rebuilt library
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu_r2.log 2>&1
echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2.log
# 3. time the transforms and the sampler at config-2 size
timeout 300 python scripts/time_transforms.py > gpurun_out/time_transforms.log 2>&1
echo "time_transforms rc=$?"; cat gpurun_out/time_transforms.log
# 4. A/B of the index-prefetch variants of the fused kernel (bit-identity is checked by the sweep itself)
VARIANTS=0,6,7,8,9,0 timeout 900 python scripts/sweep_variants.py > gpurun_out/sweep_prefetch.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/sweep_prefetch.log
# 5. one ncu capture of every kernel family (about 40 replays each at N = 2 M, E = 20 M), summarised into profiles/
timeout 1500 ncu --set full --clock-control none --import-source on -f -o gpurun_out/all_kernels \
    --kernel-name-base demangled -k regex:'gnnb|tc::|tcw::' python scripts/run_all_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu all kernels rc=$?"; tail -n 2 gpurun_out/ncu_all.log
python scripts/ncu_summarize.py gpurun_out/all_kernels.ncu-rep > gpurun_out/r2_all_kernels.md 2>/dev/null && head -n 30 gpurun_out/r2_all_kernels.md
