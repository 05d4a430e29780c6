#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  timeout 900 env "$@" python bench.py --steps 10 --warmup 3 --no-fits --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "bench_${tag}_exit=$?" | tee -a gpurun_out/status.txt
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("$tag:", round(d["value"],2), "evals/s", round(d["ms_per_step"],3), "ms  dense", round(d["phases_ms"]["potrf"],3), "lml", d["lml"])
P
  tail -3 gpurun_out/bench_$tag.err
}
run v2 CS_PLAN=2
run v2_prio_b CS_PLAN=2 CS_PRIO=2,0,1,5,4
run v2_prio_f CS_PLAN=2 CS_PRIO=1,0,2,5,3
timeout 300 python tools/timeline.py --inv-pipeline 1 --window ${WINDOW:-5.0,6.2} --out gpurun_out/timeline_pipe.csv > gpurun_out/timeline_pipe.txt 2>&1
echo "timeline_exit=$?" | tee -a gpurun_out/status.txt
