#!/bin/bash
# ncu launch list (gpu__time_duration only) of the default bench command without the fits / cpu / config legs
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
B="python bench.py --steps 2 --warmup 1 --no-fits --no-cpu --no-configs"   # the SM partition switches itself off under ncu (green-context launches cannot be profiled)
timeout 600 $B > gpurun_out/plain.log 2>&1
echo "plain_exit=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu_list_exit=$?"
