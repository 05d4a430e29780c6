"""The headline configuration (n=8192, p=25, GP and Student-t) and a ragged n=4000 case held to numbers produced
by the reference's OWN C++ (tests/golden/golden_ref_8192.npz, made by tests/golden/make_golden_ref_8192.py from
oracle/_ref = /root/reference/src/kernel_{GP,TP}_SE_cpp.cpp compiled unmodified, host LAPACK).

These are the only sizes at which the plans that exist for padded n >= 4096 run (green-context SM partition, look-ahead of
depth three, tail share of the chain partition): both the default plan (what bench.py times) and the sequential plan are
compared -- LML 1e-9, gradient 1e-8, mu 1e-8, K^-1 rows and diagonal 1e-9 (BASELINE.json tolerances).
Reference: src/kernel_GP_SE_cpp.cpp:143-202, src/kernel_TP_SE_cpp.cpp:100-163.
"""
import os

import numpy as np
import pytest

from causalstump_b200 import _lib
from tests.golden import make_golden_ref_8192 as mk

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref_8192.npz")
PLANS = {"default": dict(inv_pipeline=1, partition=0), "sequential": dict(inv_pipeline=0, partition=-1)}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("plan", ["default", "sequential"])
@pytest.mark.parametrize("name,kind,n,weights", mk.CASES)
def test_headline_size_against_the_compiled_reference(gpulib, gold, name, kind, n, weights, plan):
    y, X, z, w, par = mk.case_inputs(kind, n, weights)
    assert mk.digest(y, X, z, w, par) == str(gold[name + "/digest"]), "input generator drifted: regenerate the fixture"
    f = _lib.Fit(gpulib, _lib.CS_GP if kind == "GP" else _lib.CS_TP, y, X, z, w, par, **PLANS[plan])
    try:
        g, st, mu = f.eval()
        gref, sref, mref = gold[name + "/grad"], gold[name + "/stats"], float(gold[name + "/mu_sol"])
        gerr = float(np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))))
        assert abs(st[1] - sref[1]) <= 1e-9 * abs(sref[1]), ("LML", st[1], sref[1])
        assert gerr <= 1e-8, ("gradient", gerr)
        assert abs(st[0] - sref[0]) <= 1e-7 * abs(sref[0]), ("squared residual", st[0], sref[0])
        assert abs(mu - mref) <= 1e-8 * abs(mref), ("mu", mu, mref)
        inv = f.get_matrices(want=("invK",))["invK"]
        rows = gold[name + "/rows"]
        scale = float(np.max(np.abs(gold[name + "/invK_diag"])))
        assert np.max(np.abs(np.diag(inv) - gold[name + "/invK_diag"])) <= 1e-9 * scale
        assert np.max(np.abs(inv[rows] - gold[name + "/invK_rows"])) <= 1e-9 * scale
        assert np.array_equal(inv, inv.T)
    finally:
        f.close()


@pytest.mark.parametrize("mode", ["nccl_1x1", "sim_2x2", "sim_2x4"])
def test_distributed_engine_against_the_compiled_reference(gpulib, gold, mode):
    """Config 5's code path (csrc/dist.cuh: block-cyclic tables, three right-looking sweeps, panel sharing) at the largest
    size the compiled reference could produce: n=8192, p=25.  nccl_1x1 = a real NCCL communicator (ncclCommInitRank +
    ncclCommSplit) with one rank; sim_PxQ = all ranks of the grid in this process with device-copy collectives (the same
    driver object).  The multi-process NCCL run is test_nccl_ranks_in_processes below and `bench.py --gpus N` (config5)."""
    y, X, z, w, par = mk.case_inputs("GP", 8192, False)
    assert mk.digest(y, X, z, w, par) == str(gold["gp8192/digest"])
    if mode == "nccl_1x1":
        import torch  # noqa: F401  (loads the NCCL that cs_dist binds to by soname, as bench.py does)
        f = _lib.DistFit(gpulib, 0, y, X, z, w, par, 0, 1, 1, 1, gpulib.dist_unique_id())
        try:
            g, st, mu, _ = f.eval()
        finally:
            f.close()
    else:
        P, Q = int(mode[4]), int(mode[6])
        g, st, mu = gpulib.dist_sim_eval(0, y, X, z, w, par, P, Q, 128)
    gref, sref, mref = gold["gp8192/grad"], gold["gp8192/stats"], float(gold["gp8192/mu_sol"])
    assert abs(st[1] - sref[1]) <= 1e-9 * abs(sref[1])
    assert float(np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref))))) <= 1e-8
    assert abs(mu - mref) <= 1e-8 * abs(mref) and abs(st[0] - sref[0]) <= 1e-7 * abs(sref[0])


def test_nccl_ranks_in_processes(gpulib):
    """Two real ranks (one process per GPU, torchrun, NCCL over NVLink) through tools/dist_check.py: parity with the oracle at
    n=1500 on the 2x1 grid and agreement of all ranks.  Needs two visible GPUs; on a one-GPU box the NCCL path is covered by
    nccl_1x1 above and by the driver's `bench.py --gpus N` runs."""
    import json
    import subprocess
    import sys
    if gpulib.dll.cs_device_count() < 2:
        pytest.skip("one GPU visible: multi-process NCCL run not possible here")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "dist_check.py"), "--size", "4096", "--steps", "1", "--paritysize", "1500"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["parity"]["grad_rel"] <= 1e-8 and out["parity"]["lml_rel"] <= 1e-9 and out["parity"]["ranks_agree"] and out["finite"]
