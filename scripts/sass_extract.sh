#!/bin/bash
# SASS evidence for profiles/: full listing of the default fused kernel, and the mnemonic histograms that prove the
# Blackwell paths (UTC*MMA / LDTM = tcgen05, UBLKCP / UTMALDG = TMA).  Run after `make`.
set -e
cd "$(dirname "$0")/.."
B=graphneuralnetworks.jl_b200/build
O=profiles/sass
mkdir -p $O
hist() { grep -oE '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+' | sed -E 's/.*\s([A-Z0-9_.]+)$/\1/' | sed -E 's/\..*//' | sort | uniq -c | sort -rn; }
fn() { cuobjdump -sass "$1" | awk -v pat="$2" '/Function : /{f = ($0 ~ pat)} f' | grep -v '^\s*/\* 0x'; }
fn $B/seglean.o 'seg_lean_kernelILi1ELi1ELb0ELi0ELi0E' > $O/r2_seg_lean_kernel_D128_sum_stream.sass
fn $B/seglean.o 'seg_lean_kernelILi1ELi1ELb0ELi0ELi0E' | hist > $O/r2_seg_lean_kernel_D128_sum_stream.hist
fn $B/segbulk.o 'seg_reduce_bulk_kernelILi1ELi2ELi3ELb0E' | hist > $O/r2_seg_reduce_bulk_kernel_D128.hist
fn $B/dense_tc.o 'linear_tf32x3_kernel' | hist > $O/r2_linear_tf32x3_kernel.hist
fn $B/dense_tc.o 'dw_tf32x3_kernel' | hist > $O/r2_dw_tf32x3_kernel.hist
ls -la $O
fn $B/dense_tc.o 'linear_wide_tf32x3_kernel' | hist > $O/r2_linear_wide_tf32x3_kernel.hist
fn $B/seglean.o 'maxmin_bwd_lean_kernelILi1ELb0E' | hist > $O/r2_maxmin_bwd_lean_kernel_D128.hist
ls -la $O
