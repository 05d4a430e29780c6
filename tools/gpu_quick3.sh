#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for v in 0 1 0 1; do
if [ $v = 1 ]; then export CS_NO_GRAM_SPLIT=1; else unset CS_NO_GRAM_SPLIT; fi
timeout 900 python bench.py --steps 20 --warmup 3 --no-fits --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("nosplit=$v", round(d["value"],2), "evals/s", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()})
P
done
