#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (ours + reference arm), fits scaling, timeline of the dense plan.
# ncu evidence: tools/gpu_profile.sh.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu" | tee gpurun_out/status.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke_exit=$?" | tee -a gpurun_out/status.txt
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?" | tee -a gpurun_out/status.txt
tail -c 5000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench_ref_exit=$?" | tee -a gpurun_out/status.txt
  tail -c 1500 gpurun_out/bench_ref.json
fi
timeout 600 python tools/fits_scaling.py 64 128 256 --batch=32 > gpurun_out/fits_scaling.txt 2>&1
echo "fits_scaling_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/fits_scaling.txt
timeout 300 python tools/batch_phases.py > gpurun_out/batch_phases.txt 2>&1
cat gpurun_out/batch_phases.txt
timeout 300 python tools/timeline.py --inv-pipeline 1 --out gpurun_out/timeline_pipe.csv > gpurun_out/timeline_pipe.txt 2>&1
echo "timeline_exit=$?" | tee -a gpurun_out/status.txt
head -12 gpurun_out/timeline_pipe.txt
cat gpurun_out/status.txt
