#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/fits_scaling.py 4 16 32 64 128 > gpurun_out/fits_scaling.txt 2>&1
echo "fits_scaling_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/fits_scaling.txt
