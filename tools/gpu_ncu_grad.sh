#!/bin/bash
# one ncu --set full capture (with source) of the tensor-path trace-gradient kernel in the sequential-plan bench command
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
S="python bench.py --steps 1 --warmup 1 --no-fits --no-cpu --no-configs --inv-pipeline 0"
timeout 600 $S > gpurun_out/plain_seq.log 2>&1
echo "plain_seq_exit=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k "regex:${1:-grad_trace_mma_kernel}" -s 1 -c 1 -o gpurun_out/prof_${2:-grad} -f $S > gpurun_out/ncu_full_${2:-grad}.log 2>&1
echo "ncu_exit=$?"
