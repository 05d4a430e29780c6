#!/bin/bash
# split leaf (factor on the chain, inverse on S8) + triangular solve: chain partition size sweep below the headline size
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for sms in 8 16 24 32; do
  export CS_CRIT_SMS=$sms
  echo "== chain partition of $sms SMs"
  for n in 4096 6144 8192; do timeout 200 python tools/big_single.py $n 2>&1 | tail -1; done
done
