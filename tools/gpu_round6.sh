#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -6 gpurun_out/pytest_gpu.log
for part in 0 -1; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-fits --no-cpu --partition $part > gpurun_out/bench_part$part.json 2> gpurun_out/bench_part$part.err
echo "bench_part${part}_exit=$?" | tee -a gpurun_out/status.txt
python - <<P
import json
d=json.loads(open("gpurun_out/bench_part$part.json").read().strip().splitlines()[-1])
print("partition $part:", d["value"], "evals/s", d["ms_per_step"], "ms", d["phases_ms"])
P
tail -3 gpurun_out/bench_part$part.err
done
timeout 300 python tools/timeline.py --inv-pipeline 1 --out gpurun_out/timeline_pipe.csv > gpurun_out/timeline_pipe.txt 2>&1
echo "timeline_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/timeline_pipe.txt
