#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/fits_scaling.py 128 --batch=32 --tile=128 2>&1 | tail -2
timeout 600 env CS_PANEL=2 python tools/fits_scaling.py 128 --batch=32 2>&1 | tail -1
timeout 600 env CS_PANEL=6 python tools/fits_scaling.py 128 --batch=32 2>&1 | tail -1
timeout 600 env CS_PLAN=1 python tools/fits_scaling.py 128 --batch=32 2>&1 | tail -1
timeout 600 python tools/fits_scaling.py 256 --batch=32 2>&1 | tail -1
