#!/bin/bash
# ncu evidence for profiles/: reports stay on the box (gpurun_out/ is capped at 64 MiB); the raw-metric CSVs, the summary
# table and the launch list come back.  Small reports (< 12 MB) are copied too.
set -u
mkdir -p gpurun_out /tmp/rep
GNNB_BENCH_PARTITIONED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29520 bench.py --config 2 --nodes 2000000 --edges 20000000 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_partitioned_1rank.log 2>&1
echo "partitioned path on one rank rc=$?"; tail -n 1 gpurun_out/bench_partitioned_1rank.log | python -c "import sys, json; l = sys.stdin.read(); print(json.loads(l)['parity_rel_err'] if l.startswith('{') else l[-1500:])"
grep -B2 -A12 Traceback gpurun_out/bench_partitioned_1rank.log | head -n 40
S=$SECONDS
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_final_reference.log 2>&1
echo "reference arm rc=$? elapsed=$((SECONDS - S))s"; tail -n 1 gpurun_out/bench_final_reference.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_gcn_step.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_step.log 2>&1
echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/prof_gat_lean_r2 --kernel-name-base demangled -k regex:'gat_(fwd|bwd)_lean' -s 2 -c 2 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_gat.log 2>&1
echo "ncu gat rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/prof_lean_mean_c4_r2 --kernel-name-base demangled -k regex:'seg_lean_kernel<1, 0, 0, 0, 1>' -s 2 -c 1 python bench.py --config 4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/prof_tc_r2 --kernel-name-base demangled -k regex:'linear_tf32x3|dw_tf32x3' -c 2 python scripts/profile_linear.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc rc=$?"
timeout 1200 ncu --set full --clock-control none -f -o /tmp/rep/all_kernels_r2 --kernel-name-base demangled -k regex:'gnnb|tc::|tcw::' python scripts/run_all_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu all kernels rc=$?"; tail -n 2 gpurun_out/ncu_all.log | cut -c1-200
for r in prof_gat_lean_r2 prof_lean_mean_c4_r2 prof_tc_r2; do
  [ -f /tmp/rep/$r.ncu-rep ] && python scripts/ncu_raw_extract.py /tmp/rep/$r.ncu-rep > gpurun_out/${r}_ncu_raw.csv
done
[ -f /tmp/rep/all_kernels_r2.ncu-rep ] && python scripts/ncu_summarize.py /tmp/rep/all_kernels_r2.ncu-rep > gpurun_out/all_kernels_r2.md
[ -f /tmp/rep/all_kernels_r2.ncu-rep ] && ncu -i /tmp/rep/all_kernels_r2.ncu-rep --page raw --csv > gpurun_out/all_kernels_r2_raw.csv
ls -la /tmp/rep
for f in /tmp/rep/*.ncu-rep; do
  [ "$(stat -c %s "$f")" -lt 12000000 ] && cp "$f" gpurun_out/
done
du -sh gpurun_out
