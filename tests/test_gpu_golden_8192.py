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
