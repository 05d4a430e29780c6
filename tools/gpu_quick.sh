#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
