#!/bin/bash
# tensor-path Gram / trace-gradient kernels (kernels_mma.cuh) vs the element-wise ones (CS_GRAD_OLD=1)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
for v in MMA OLD; do
  if [ $v = OLD ]; then export CS_GRAD_OLD=1; else unset CS_GRAD_OLD; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-fits --no-cpu --no-configs > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
print("$v", {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["phases_ms_sequential_plan"], d["lml"])
PY
  timeout 300 python tools/batch_phases.py > gpurun_out/batch_phases_$v.txt 2>&1; grep "tile= 64" gpurun_out/batch_phases_$v.txt
  timeout 300 python tools/fits_scaling.py 256 --batch=32 --device > gpurun_out/fits_$v.txt 2>&1; cat gpurun_out/fits_$v.txt
done
