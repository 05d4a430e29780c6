#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu" | tee gpurun_out/status.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee -a gpurun_out/status.txt
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== dist_check 1 rank n=8192"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py --n 8192 --steps 2 --parity-n 1500 > gpurun_out/dist1_8192.json 2> gpurun_out/dist1_8192.err
echo "dist1_8192_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/dist1_8192.json; tail -5 gpurun_out/dist1_8192.err
echo "== dist_check 1 rank n=32768"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py --n 32768 --steps 1 --parity-n 0 > gpurun_out/dist1_32768.json 2> gpurun_out/dist1_32768.err
echo "dist1_32768_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/dist1_32768.json; tail -5 gpurun_out/dist1_32768.err
cat gpurun_out/status.txt
