#!/bin/bash
# ncu evidence for the current build: launch list of the default bench command and --set full captures of the
# dominant kernel (K^-1 = X^T X as one launch, sequential plan), a K=512 trailing-update launch and the leaf.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
B="python bench.py --steps 2 --warmup 1 --no-fits --no-cpu"   # the SM partition switches itself off under ncu (green-context launches cannot be profiled)
timeout 600 $B > gpurun_out/plain.log 2>&1
echo "plain_exit=$?" | tee gpurun_out/status.txt
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu_list_exit=$?" | tee -a gpurun_out/status.txt
S="python bench.py --steps 1 --warmup 1 --no-fits --no-cpu --inv-pipeline 0"
timeout 600 $S > gpurun_out/plain_seq.log 2>&1
echo "plain_seq_exit=$?" | tee -a gpurun_out/status.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:gemm_kernel<.int.64, .int.2, .int.2, .bool.1, .bool.1' -s 2 -c 1 -o gpurun_out/prof_xtx -f $S > gpurun_out/ncu_full_xtx.log 2>&1
echo "ncu_xtx_exit=$?" | tee -a gpurun_out/status.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:grad_trace_mma_kernel' -s 1 -c 1 -o gpurun_out/prof_grad -f $S > gpurun_out/ncu_full_grad.log 2>&1
echo "ncu_grad_exit=$?" | tee -a gpurun_out/status.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:gram_train_mma_kernel' -s 1 -c 1 -o gpurun_out/prof_gram -f $S > gpurun_out/ncu_full_gram.log 2>&1
echo "ncu_gram_exit=$?" | tee -a gpurun_out/status.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k 'regex:leaf_kernel' -s 10 -c 1 -o gpurun_out/prof_leaf -f $S > gpurun_out/ncu_full_leaf.log 2>&1
echo "ncu_leaf_exit=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/status.txt
