#!/bin/bash
# gradient / Gram kernel changes: parity tests, phase times at n=8192 and in a batch, fits
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "bench_exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print("phases_seq", d["phases_ms_sequential_plan"])
print("fits", d["fits"]["value"])
PY
timeout 300 python tools/batch_phases.py 2>&1 | tail -4
