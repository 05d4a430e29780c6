#!/bin/bash
# device-side WHILE loop (one graph launch per cs_fit_run) vs the chunked host loop: GPU tests, then fits/hour both ways
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_exit=$?" | tee gpurun_out/status.txt
tail -6 gpurun_out/pytest_gpu.log
for v in WHILE HOSTLOOP; do
  if [ $v = HOSTLOOP ]; then export CS_NO_WHILE=1; else unset CS_NO_WHILE; fi
  timeout 300 python tools/fits_scaling.py 64 256 512 --batch=32 --device > gpurun_out/fits_$v.txt 2>&1
  echo "== $v"; cat gpurun_out/fits_$v.txt
done
unset CS_NO_WHILE
timeout 120 python - <<'PY'
import time, numpy as np, bench
from causalstump_b200 import _lib
lib = _lib.load()
for n in (120, 747):
    inp = bench.make_inputs(n, 25 if n > 200 else 1, 1)
    f = _lib.Fit(lib, 0, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"])
    f.run(50, 0.0)
    t0 = time.perf_counter(); it, tr = f.run(400, 0.0); dt = time.perf_counter() - t0
    print("single fit n=%d: %d iterations in %.3f s = %.1f us per iteration (device loop)" % (n, it, dt, 1e6 * dt / it))
    f.close()
PY
cat gpurun_out/status.txt
