#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -5 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-fits --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print(round(d["value"],2), "evals/s", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()})
P
done
tail -3 gpurun_out/bench_q.err
