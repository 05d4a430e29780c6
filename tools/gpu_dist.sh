#!/bin/bash
# distributed (config 5) check over N ranks: bash tools/gpu_dist.sh N SIZE STEPS
set -u
N=${1:-2}; SIZE=${2:-32768}; STEPS=${3:-2}; PAR=${4:-1500}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.csv 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   tools/dist_check.py --size $SIZE --steps $STEPS --paritysize $PAR > gpurun_out/dist${N}_${SIZE}.json 2> gpurun_out/dist${N}_${SIZE}.err
echo "dist${N}_${SIZE}_exit=$?"
cat gpurun_out/dist${N}_${SIZE}.json; tail -8 gpurun_out/dist${N}_${SIZE}.err
