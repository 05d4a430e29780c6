#!/bin/bash
# A/B of the round-2 changes on one B200: GPU tests, then the batched small-n iteration with the left-looking vs the
# right-looking Cholesky plan, then the n=8192 evaluation with and without TMA-staged covariate tiles.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -8 gpurun_out/pytest_gpu.log
for v in LL RIGHT; do
  if [ $v = RIGHT ]; then export CS_BATCH_RIGHT=1; else unset CS_BATCH_RIGHT; fi
  timeout 300 python tools/batch_phases.py > gpurun_out/batch_phases_$v.txt 2>&1
  echo "== batch_phases $v"; grep "tile= 64" gpurun_out/batch_phases_$v.txt
  timeout 300 python tools/fits_scaling.py 256 --batch=32 > gpurun_out/fits_scaling_$v.txt 2>&1
  echo "== fits_scaling $v"; cat gpurun_out/fits_scaling_$v.txt
done
unset CS_BATCH_RIGHT
for v in TMA NOTMA; do
  if [ $v = NOTMA ]; then export CS_NO_TMA=1; else unset CS_NO_TMA; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-fits --no-cpu --no-configs > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "== bench $v exit=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["phases_ms_sequential_plan"], d["roofline"]["frac"])
PY
done
cat gpurun_out/status.txt
