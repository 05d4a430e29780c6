#!/bin/bash
# round 2 soak: GPU suite x3, randomised sweeps vs the oracle and vs the compiled reference (p up to 100), sim / batched tests x3
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
fail=0
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/soak2_$i.log 2>&1 || fail=$((fail+1))
  tail -1 gpurun_out/soak2_$i.log
done
timeout 900 python tools/parity_sweep.py 20261020 60 > gpurun_out/r02_parity_sweep_oracle.txt 2>&1 || fail=$((fail+1))
tail -3 gpurun_out/r02_parity_sweep_oracle.txt | cut -c1-300
timeout 900 python tools/parity_sweep.py 20261021 40 --ref > gpurun_out/r02_parity_sweep_compiled_reference.txt 2>&1 || fail=$((fail+1))
tail -3 gpurun_out/r02_parity_sweep_compiled_reference.txt | cut -c1-300
for i in 1 2 3; do
  timeout 600 python -m pytest tests/test_sim_driver.py tests/test_gpu_parity.py -m gpu -x -q -k "sim or batched or failing or concurrent or chunking" > gpurun_out/soak2_b$i.log 2>&1 || fail=$((fail+1))
  tail -1 gpurun_out/soak2_b$i.log
done
echo "soak failures: $fail"
