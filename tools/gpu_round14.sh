#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -15 gpurun_out/pytest_gpu.log
for b in 1 8 16 32 64; do
timeout 600 python tools/fits_scaling.py 64 128 --batch=$b > gpurun_out/fits_scaling_b$b.txt 2>&1
echo "fits_scaling_b${b}_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/fits_scaling_b$b.txt | tail -3
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --fits 64 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?" | tee -a gpurun_out/status.txt
python - <<P
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(round(d["value"],2), "evals/s", d["fits"])
P
tail -3 gpurun_out/bench.err
