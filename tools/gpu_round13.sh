#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
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
run w2 CS_PANEL=2
run w3 CS_PANEL=3
run w6 CS_PANEL=6
run w8 CS_PANEL=8
run w8_v1 CS_PANEL=8 CS_PLAN=1
