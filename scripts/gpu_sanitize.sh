#!/bin/bash
# the kernels added late in round 2 under compute-sanitizer: TOOL=memcheck (default) | racecheck | synccheck
set -u
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
timeout 60 compute-sanitizer --tool $TOOL --error-exitcode 3 python scripts/sanitize_new_kernels.py > gpurun_out/sanitize_new_kernels_$TOOL.log 2>&1
echo "sanitizer ($TOOL) rc=$?"; tail -n 14 gpurun_out/sanitize_new_kernels_$TOOL.log | cut -c1-300
