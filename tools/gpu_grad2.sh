#!/bin/bash
# quick A/B of the gradient kernel: phase times at n=8192 and in a batch for a few CS_GRAD_CTAS values
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_8192.py -m gpu -x -q -k "not nccl and not distributed" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?"; tail -2 gpurun_out/pytest_gpu.log
for c in ${GRAD_CTAS_LIST:-2 1}; do
  true
  echo "== CS_GRAD_CTAS=$c"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --no-fits > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, "grad", d["phases_ms_sequential_plan"]["grad"], "gram", d["phases_ms_sequential_plan"]["gram"])
PY
  timeout 300 python tools/batch_phases.py 2>&1 | grep "B=32 tile= 64\|B=64 tile= 64"
done
