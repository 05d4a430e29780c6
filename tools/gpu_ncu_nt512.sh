#!/bin/bash
# ncu --set full of ONE K = 512 trailing-update launch (NT, 64 x 64 tiles, the shipped 2-stage / 4-CTA configuration) of the microbenchmark
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 ./tools/gemm_bench > gpurun_out/gemm_bench_plain.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:gemm_kernel<.int.64, .int.2, .int.2, .bool.0, .bool.0, .int.16, .int.2, .int.4' -s 1 -c 1 -o gpurun_out/prof_nt512 -f ./tools/gemm_bench > gpurun_out/ncu_nt512.log 2>&1
echo "exit=$?"; tail -5 gpurun_out/ncu_nt512.log
