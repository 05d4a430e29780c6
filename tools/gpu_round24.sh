#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/fits_scaling.py 128 256 --batch=32 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum -c 1 python -c "
import ctypes, os
print([l.split()[-1] for l in open('/proc/self/maps') if any(k in l.lower() for k in ('nsight','inject','ncu','perfworks','cupti'))][:12])
print({k:v for k,v in os.environ.items() if any(s in k.upper() for s in ('INJECT','NSIGHT','NCU','PRELOAD','CUPTI'))})
" 2>&1 | tail -4
