#!/bin/bash
# two GPUs: the multi-process NCCL test, then bench.py --gpus 2 as the driver launches it (config4 + config5 objects)
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_golden_8192.py -m gpu -q -k "nccl" > gpurun_out/pytest_nccl_n$N.log 2>&1
echo "pytest_nccl_exit=$?" | tee gpurun_out/status_n$N.txt
tail -4 gpurun_out/pytest_nccl_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench_n${N}_exit=$?" | tee -a gpurun_out/status_n$N.txt
tail -c 6000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
