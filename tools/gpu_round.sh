#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (ours + reference), then the ncu launch
# list and one --set full capture of the dominant kernel -- each ncu pass only after the
# same command exited 0 without ncu.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
if [ -x tools/gemm_bench ] && [ "${SKIP_GEMM_BENCH:-0}" != "1" ]; then
  timeout 600 tools/gemm_bench > gpurun_out/gemm_bench.txt 2>&1
  echo "gemm_bench_exit=$?"
  cat gpurun_out/gemm_bench.txt
fi
echo "== pytest -m gpu" | tee gpurun_out/status.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee -a gpurun_out/status.txt
tail -25 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke_exit=$?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
BE=$?
echo "bench_exit=$BE" | tee -a gpurun_out/status.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench_ref_exit=$?" | tee -a gpurun_out/status.txt
  cat gpurun_out/bench_ref.json
fi
if [ "$BE" = "0" ] && [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 python bench.py --steps 2 --warmup 1 --no-fits --no-cpu > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-fits --no-cpu > gpurun_out/ncu_list.log 2>&1
  echo "ncu_list_exit=$?" | tee -a gpurun_out/status.txt
  echo "== ncu full (dominant kernel)"
  timeout 600 python bench.py --steps 1 --warmup 1 --no-fits --no-cpu > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k 'regex:gemm_kernel<.int.64, .int.2, .int.2, .bool.1, .bool.1' -s 1 -c 1 \
      -o gpurun_out/prof_xtx -f python bench.py --steps 1 --warmup 1 --no-fits --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "ncu_full_exit=$?" | tee -a gpurun_out/status.txt
  timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k 'regex:leaf_kernel' -s 10 -c 1 \
      -o gpurun_out/prof_leaf -f python bench.py --steps 1 --warmup 1 --no-fits --no-cpu > gpurun_out/ncu_full2.log 2>&1
  echo "ncu_full2_exit=$?" | tee -a gpurun_out/status.txt
fi
cat gpurun_out/status.txt
