#!/bin/bash
# final single-GPU validation of the round: whole GPU suite, smoke, the driver's bench commands, ncu evidence for profiles/
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest -m gpu rc=$?"; tail -n 6 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_1gpu.log 2>&1
echo "bench rc=$?"; tail -n 1 gpurun_out/bench_final_1gpu.log | cut -c1-600
/usr/bin/time -v timeout 1200 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_final_reference.log 2> gpurun_out/bench_final_reference.time
echo "reference arm rc=$?"; tail -n 1 gpurun_out/bench_final_reference.log | cut -c1-900; grep "Elapsed\|Maximum resident" gpurun_out/bench_final_reference.time
# ncu: launch list of one bench step, and --set full captures of the kernels the other configs' roofline lines name
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_gcn_step.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_step.log 2>&1
echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_gat_lean_r2 --kernel-name-base demangled -k regex:'gat_(fwd|bwd)_lean' -s 2 -c 2 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_gat.log 2>&1
echo "ncu gat rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_lean_mean_c4_r2 --kernel-name-base demangled -k regex:'seg_lean_kernel<1, 0, 0, 0, 1>' -s 2 -c 1 python bench.py --config 4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_tc_r2 --kernel-name-base demangled -k regex:'linear_tf32x3|dw_tf32x3' -c 2 python scripts/profile_linear.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc rc=$?"
# one ncu --set full pass over every kernel family (N = 2 M, E = 20 M), summarised into a table
timeout 1200 ncu --set full --clock-control none -f -o gpurun_out/all_kernels_r2 --kernel-name-base demangled -k regex:'gnnb|tc::|tcw::' python scripts/run_all_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu all kernels rc=$?"; tail -n 2 gpurun_out/ncu_all.log | cut -c1-200
