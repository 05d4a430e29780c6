#!/bin/bash
# N GPUs: bench.py --gpus N exactly as the driver launches it (config4 + config5 objects)
set -u
N=${1:-8}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.csv 2>&1
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench_n${N}_exit=$? after ${SECONDS}s" | tee gpurun_out/status_n$N.txt
tail -c 7000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
