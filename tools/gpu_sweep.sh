cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 python tools/batch_phases.py 2>&1 | grep "B= 1 "
CS_PANEL=4 timeout 500 python tools/batch_phases.py 2>&1 | grep "B= 1 "
timeout 300 python tools/parity_sweep.py 5 40 2>&1 | tail -1 | cut -c1-200
timeout 300 python - <<'P'
import time, sys
sys.path.insert(0, '.')
import numpy as np, bench
from causalstump_b200 import _lib
lib = _lib.load()
for n in (120, 400, 672, 747):
    inp = bench.make_inputs(n, 25 if n > 120 else 1, 0)
    f = _lib.Fit(lib, 0, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"])
    f.run(50, 0.0)
    t0 = time.perf_counter(); it, tr = f.run(300, 0.0); dt = time.perf_counter() - t0
    print("single fit n=%d: %.3f ms per optimizer iteration (300 iterations, device loop incl. polling)" % (n, 1e3 * dt / 300))
    f.close()
P
