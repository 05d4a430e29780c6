#!/bin/bash
# 2-GPU run: bench as the driver launches it (torchrun, one rank per GPU), reference arm, then the ncu launch list
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.csv 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench_n2_exit=$?" | tee gpurun_out/status.txt
tail -c 3000 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench_exit=$?" | tee -a gpurun_out/status.txt
tail -c 6000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench_ref_exit=$?" | tee -a gpurun_out/status.txt
tail -c 1500 gpurun_out/bench_ref.json
B="python bench.py --steps 2 --warmup 1 --no-fits --no-cpu --partition -1"
timeout 600 $B > gpurun_out/plain.log 2>&1
echo "plain_exit=$?" | tee -a gpurun_out/status.txt
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu_list_exit=$?" | tee -a gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke_exit=$?" | tee -a gpurun_out/status.txt
tail -2 gpurun_out/smoke.log
