#!/bin/bash
# per-launch FP64 tensor-pipe utilisation of every GEMM launch of one evaluation (the Cholesky / inverse updates)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-fits --no-cpu --no-configs"
timeout 600 $B > gpurun_out/plain_g.log 2>&1; echo "plain_exit=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --kernel-name-base demangled -k 'regex:gemm_kernel' -s 1400 -c 460 --csv --log-file gpurun_out/gemm_launches.csv $B > gpurun_out/ncu_gemms.log 2>&1
echo "ncu_exit=$?"
