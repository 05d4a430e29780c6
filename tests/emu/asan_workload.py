"""Workload of tests/test_emu_asan.py: the kernels and launch plans on the SIMT emulator built with
AddressSanitizer (device memory = host heap there, so out-of-bounds tile / panel accesses are caught)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from causalstump_b200 import _lib  # noqa: E402
from oracle import causalstump_oracle as o  # noqa: E402
from tests import _cases as cs  # noqa: E402

lib = _lib.CsLib(sys.argv[1])
cs.check_eval(lib, "GP", 130, 5)
cs.check_eval(lib, "TP", 65, 3, tile=64, nu=20.0)
cs.check_fit(lib, "GP", 90, 3, maxiter=3, n2=17)
Ps = [cs.make_problem("GP", 100, 4, replicate=r) for r in range(2)]
fb = _lib.FitBatch(lib, 0, [(P["y"], P["X"], P["z"], None, o.pack_params(P["par"])) for P in Ps], chunk=2)
fb.run(3, 1e-12)
fb.select(1).treat(Ps[1]["X"][:20])
fb.close()
for n, tile in ((330, 64), (300, 128)):
    P = cs.make_problem("GP", n, 3, weights=True)
    f = _lib.Fit(lib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=tile, inv_pipeline=1)
    f.eval()
    f.close()
P = cs.make_problem("GP", 200, 4)
lib.dist_sim_eval(0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), 2, 2, 64)
print("asan workload done")
