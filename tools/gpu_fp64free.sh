#!/bin/bash
# effect of taking the fp64 divide / sqrt / multiplies out of the GEMM tile prologue and epilogue
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 ./tools/gemm_bench > gpurun_out/gemm_bench_r02.txt 2>&1
grep -E "peak|==|64 2x2 kc16 s[23]" gpurun_out/gemm_bench_r02.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-fits --no-cpu --no-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["phases_ms_sequential_plan"])
PY
timeout 300 python tools/batch_phases.py > gpurun_out/batch_phases_r02.txt 2>&1; grep "tile= 64" gpurun_out/batch_phases_r02.txt
timeout 300 python tools/batch_timeline.py 64 672 > gpurun_out/batch_timeline.txt 2>&1; cat gpurun_out/batch_timeline.txt
timeout 300 python tools/fits_scaling.py 256 512 --batch=32 --device > gpurun_out/fits_r02.txt 2>&1; cat gpurun_out/fits_r02.txt
