#!/bin/bash
# fused distributed sweep on N GPUs: dist tests + config5 through tools/dist_check.py, old two-sweep driver for comparison
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_golden_8192.py tests/test_gpu_parity.py -m gpu -q -k "distributed or nccl" > gpurun_out/pytest_dist_n$N.log 2>&1
echo "pytest_dist_exit=$?"; tail -4 gpurun_out/pytest_dist_n$N.log
for v in FUSED TWO; do
  if [ $v = TWO ]; then export CS_DIST_TWO_SWEEPS=1; else unset CS_DIST_TWO_SWEEPS; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     tools/dist_check.py --size 32768 --steps 2 --paritysize 1500 > gpurun_out/dist${N}_$v.json 2> gpurun_out/dist${N}_$v.err
  echo "$v exit=$?"; cat gpurun_out/dist${N}_$v.json; tail -2 gpurun_out/dist${N}_$v.err
done
