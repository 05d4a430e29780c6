#!/bin/bash
# gpurun call: GPU parity tests, full default bench, timeline of the dense plan (pipelined and sequential)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/timeline.py --inv-pipeline 1 --out gpurun_out/timeline_pipe.csv > gpurun_out/timeline_pipe.txt 2>&1
echo "timeline_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/timeline_pipe.txt
timeout 300 python tools/timeline.py --inv-pipeline 0 --out gpurun_out/timeline_seq.csv > gpurun_out/timeline_seq.txt 2>&1
cat gpurun_out/timeline_seq.txt
