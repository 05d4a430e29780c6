#!/bin/bash
# thin launches keep three stages; tail share sweep with the faster bulk kernel
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for v in default s2; do
  unset CS_GEMM_S2; [ $v = s2 ] && export CS_GEMM_S2=1
  echo "== $v"
  for n in 747 2048 4096 6144 8192; do timeout 200 python tools/big_single.py $n 2>&1 | tail -1; done
done
unset CS_GEMM_S2
for a in 1 2 3 "3,10" "3,12" "4,12"; do export CS_TAIL_A=$a; timeout 200 python tools/big_single.py 8192 2>&1 | tail -1; done
unset CS_TAIL_A
timeout 300 python tools/batch_phases.py 2>&1 | tail -2
