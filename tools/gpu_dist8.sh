#!/bin/bash
# config 5 on N GPUs: fused sweep vs two sweeps (no pytest)
set -u
N=${1:-8}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for v in FUSED TWO; do
  if [ $v = TWO ]; then export CS_DIST_TWO_SWEEPS=1; else unset CS_DIST_TWO_SWEEPS; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     tools/dist_check.py --size 32768 --steps 3 --paritysize 1500 > gpurun_out/dist${N}_$v.json 2> gpurun_out/dist${N}_$v.err
  echo "$v exit=$?"; cat gpurun_out/dist${N}_$v.json; tail -2 gpurun_out/dist${N}_$v.err
done
