#!/bin/bash
# GEMM kernel with single-thread tile setup and two stages: GPU tests, bench (no reference arm), batch phases, mid sizes, config 5 on one GPU
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "iso", d["roofline_kernel_isolated"]["achieved"])
print("phases_seq", d["phases_ms_sequential_plan"])
print("fits", d["fits"]["value"], "config4", d["config4"]["seconds"], d["config4"]["slot_utilisation"], "config5", d["config5"]["ms_per_eval"], d["config5"].get("single_gpu_engine"))
print("configs", d.get("configs_1_2_3"))
PY
timeout 300 python tools/batch_phases.py > gpurun_out/batch_phases.txt 2>&1; tail -4 gpurun_out/batch_phases.txt
for n in 747 2048 4096 6144; do timeout 200 python tools/big_single.py $n 2>&1 | tail -1; done
timeout 300 python tools/timeline.py --inv-pipeline 1 --out gpurun_out/timeline_pipe.csv > gpurun_out/timeline_pipe.txt 2>&1; head -14 gpurun_out/timeline_pipe.txt
