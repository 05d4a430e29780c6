#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-fits --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print(round(d["value"],2), "evals/s", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, d["lml"])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py --n 8192 --steps 2 --parity-n 1500 > gpurun_out/dist2_8192.json 2> gpurun_out/dist2_8192.err
echo "dist2_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/dist2_8192.json; tail -3 gpurun_out/dist2_8192.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/dist_check.py --n 32768 --steps 1 --parity-n 0 > gpurun_out/dist2_32768.json 2> gpurun_out/dist2_32768.err
echo "dist2_32768_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/dist2_32768.json; tail -3 gpurun_out/dist2_32768.err
