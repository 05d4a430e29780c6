#!/bin/bash
# repeat the GPU parity suite and the randomised sweeps to expose intermittent ordering bugs of the multi-stream plans
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
fail=0
for i in 1 2 3 4; do
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/soak_$i.log 2>&1 || fail=$((fail+1))
  tail -1 gpurun_out/soak_$i.log
done
for s in 11 22 33; do
  timeout 600 python tools/parity_sweep.py $s 40 > gpurun_out/soak_sweep_$s.log 2>&1 || fail=$((fail+1))
  tail -1 gpurun_out/soak_sweep_$s.log | cut -c1-200
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bench_plan or partition or pipelined" --count 1 > /dev/null 2>&1
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bench_plan or sm_partition or pipelined_inverse" > gpurun_out/soak_plan_$i.log 2>&1 || fail=$((fail+1))
  tail -1 gpurun_out/soak_plan_$i.log
done
echo "soak failures: $fail"
