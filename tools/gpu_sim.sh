#!/bin/bash
# device simulation driver: GPU tests of the driver, then fits/hour host-driver vs device-driver, batch sizes
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_sim_driver.py tests/test_gpu_parity.py -m gpu -q -k "sim or batched or real_ihdp" > gpurun_out/pytest_sim.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -15 gpurun_out/pytest_sim.log
for b in 32 64; do
  timeout 300 python tools/fits_scaling.py 256 512 --batch=$b > gpurun_out/fits_host_b$b.txt 2>&1
  echo "== host b=$b"; cat gpurun_out/fits_host_b$b.txt
  timeout 300 python tools/fits_scaling.py 256 512 --batch=$b --device > gpurun_out/fits_dev_b$b.txt 2>&1
  echo "== device b=$b"; cat gpurun_out/fits_dev_b$b.txt
done
cat gpurun_out/status.txt
