"""The R-side glue of the drop-in (causalstump_b200/r/src/b200_glue.cpp: the 11 registered .Call routines re-implemented on
the C ABI, plus the b200_* fast path) is compiled against oracle/shim/Rcpp.h and linked to the SIMT-emulated C-ABI library,
then called NEXT TO the reference's own compiled routines (oracle/_ref) with identical arguments: same names, same argument
lists, same numbers -- the drop-in claim at the boundary the reference actually uses, without R.  (R itself is not
installed; what is not exercised is Rcpp's real marshalling.)"""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from oracle import causalstump_oracle as o
from oracle import ref_binding as rb
from tests import _cases as cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libcs_emu.so")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    if not shutil.which("g++") or rb.build() is None:
        pytest.skip("g++ or oracle/_ref not available")
    if not os.path.exists(EMU):
        import subprocess
        subprocess.check_call(["sh", os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    out = str(tmp_path_factory.mktemp("glue") / "libglue_emu.so")
    rb.build_glue(EMU, out)
    return rb.RefLib(), rb.RefLib(out, prefix="glue")


@pytest.mark.parametrize("kind,n,p,weights", [("GP", 90, 3, True), ("TP", 70, 2, False), ("GP", 130, 25, False)])
def test_glue_routines_equal_the_reference_routines(libs, kind, n, p, weights):
    ref, glue = libs
    P = cs.make_problem(kind, n, p, weights=weights)
    par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
    k = 0 if kind == "GP" else 1
    flat = o.pack_params(par)
    la = par["lambdaa"] if k == 0 else par["lambda0"]
    a = ref.kernmat(k, X[:31], X, z[:31], z, par["Lm"], par["La"], par["lambdam"], la)
    b = glue.kernmat(k, X[:31], X, z[:31], z, par["Lm"], par["La"], par["lambdam"], la)
    for u, v in zip(a, b):
        assert cs.relm(v, u) < 1e-13
    K, Km, Ka = ref.kernmat(k, X, X, z, z, par["Lm"], par["La"], par["lambdam"], la)
    invR, invG = ref.invkernel(z, w, K, par["sigma"]), glue.invkernel(z, w, K, par["sigma"])
    assert cs.relm(invG, invR) < 1e-9
    assert abs(glue.mu_solution(y, z, invR) - ref.mu_solution(y, z, invR)) < 1e-11 * abs(ref.mu_solution(y, z, invR))
    gR, sR = ref.grad(k, y, X, z, w, K, Km, Ka, invR, flat)
    gG, sG = glue.grad(k, y, X, z, w, K, Km, Ka, invR, flat)
    assert np.max(np.abs(gG - gR) / np.maximum(np.abs(gR), 1e-6 * np.max(np.abs(gR)))) < cs.TOL_GRAD
    assert abs(sG[1] - sR[1]) < cs.TOL_LML * abs(sR[1]) and abs(sG[0] - sR[0]) < 1e-7 * abs(sR[0])
    s2R, s2G = ref.stats(k, p, y, K, invR, flat), glue.stats(k, p, y, K, invR, flat)
    assert abs(s2G[1] - s2R[1]) < cs.TOL_LML * abs(s2R[1]) and abs(s2G[0] - s2R[0]) < 1e-7 * abs(s2R[0])
    with pytest.raises(np.linalg.LinAlgError):
        glue.invkernel(np.zeros(2), np.ones(2), np.array([[1.0, 2.0], [2.0, 1.0]]), -40.0)
    glue.dll.glue_last_error.restype = C.c_char_p
    assert b"not positive definite" in glue.dll.glue_last_error()   # the R condition text of the reference


@pytest.mark.parametrize("optim", [0, 1, 2])
def test_glue_optimizers_equal_the_reference_optimizers(libs, optim):
    """Adam / Nadam / Nesterov through the glue move every list entry like the reference, including `mu` (the moment lists
    hold it, R/optimizer_classes.R:31) -- the override by the closed-form mean happens in the R class, not here."""
    ref, glue = libs
    rng = np.random.default_rng(3)
    p, kind = 4, 1
    NP = 2 * p + 5
    para, m, v = rng.normal(size=NP), 0.1 * rng.normal(size=NP), 0.1 * np.abs(rng.normal(size=NP))
    mR, vR, pR = m.copy(), v.copy(), para.copy()
    mG, vG, pG = m.copy(), v.copy(), para.copy()
    for it in range(1, 4):
        g = rng.normal(size=NP)
        b1 = 0.5 if optim == 2 else 0.9
        mR, vR, pR = ref.opt_step(optim, kind, p, it, 0.01, b1, 0.999, 1e-8, mR, vR, g, pR)
        mG, vG, pG = glue.opt_step(optim, kind, p, it, 0.01, b1, 0.999, 1e-8, mG, vG, g, pG)
        assert np.allclose(pG, pR, rtol=1e-13, atol=0) and np.allclose(mG, mR, rtol=1e-13, atol=1e-300)
        if optim != 2:
            assert np.allclose(vG, vR, rtol=1e-13, atol=1e-300)


def test_glue_fast_path_equals_the_restated_r_loop(libs):
    """b200_fit_create + b200_fit_run + b200_treat against the loop of R/main.R:79-88 driven with the
    reference's own routines (tests/golden/make_golden_ref.py::fit_loop)."""
    import importlib.util
    ref, glue = libs
    spec = importlib.util.spec_from_file_location("mgr", os.path.join(ROOT, "tests", "golden", "make_golden_ref.py"))
    mgr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgr)
    P = cs.make_problem("GP", 80, 3, weights=True)
    y, X, z, w = P["y"], np.asfortranarray(P["X"]), P["z"], P["w"]
    p0 = o.pack_params(P["par"])
    iters = 5
    parR, traceR, invK = mgr.fit_loop(ref, 0, 3, y, X, z, w, p0, "Nadam", iters, 0.01, 0.0)
    X2 = np.asfortranarray(X[:13] * 0.9)
    tmR, _, ateR, _ = mgr.predict_treat(ref, 0, 3, y, X, z, X2, parR, invK)
    it = C.c_int(0)
    trace = np.zeros((2, iters + 2), order="F")
    pout, tmap, ate = np.zeros_like(p0), np.zeros(13), C.c_double(0.0)
    f = glue.dll.glue_fit_run
    f.restype = C.c_int
    rc = f(0, 1, 80, 3, y.ctypes.data_as(dp), X.ctypes.data_as(dp), z.ctypes.data_as(dp), w.ctypes.data_as(dp), p0.ctypes.data_as(dp),
           C.c_double(0.01), C.c_double(0.0), iters, C.c_double(0.0), C.byref(it), trace.ctypes.data_as(dp), pout.ctypes.data_as(dp),
           13, X2.ctypes.data_as(dp), tmap.ctypes.data_as(dp), C.byref(ate))
    assert rc == 0 and it.value == iters
    assert cs.rel(pout, parR, 1e-9) < 1e-8 and cs.rel(trace[1, 1:], traceR[1, 1:]) < 1e-9
    assert cs.relm(tmap, tmR) < 1e-8 and abs(ate.value - ateR) < 1e-8 * max(abs(ateR), 1e-3)


def test_glue_batched_handle(libs):
    """b200_fit_create_batch / b200_fit_run_batch / b200_fit_select / b200_fit_get_params: every member equals the
    single-handle fast path on its own data."""
    ref, glue = libs
    B, n, p, iters = 3, 60, 2, 4
    Ps = [cs.make_problem("GP", n, p, replicate=r, draws=(0.4 + 0.1 * r, 0.7)) for r in range(B)]
    ys = np.ascontiguousarray(np.stack([P["y"] for P in Ps]))
    Xs = np.ascontiguousarray(np.stack([np.asfortranarray(P["X"]).ravel(order="F") for P in Ps]))
    zs = np.ascontiguousarray(np.stack([P["z"] for P in Ps]))
    p0 = np.ascontiguousarray(np.stack([o.pack_params(P["par"]) for P in Ps]))
    its = (C.c_int * B)()
    pout = np.zeros_like(p0)
    f = glue.dll.glue_fit_run_batch
    f.restype = C.c_int
    rc = f(0, 1, B, n, p, ys.ctypes.data_as(dp), Xs.ctypes.data_as(dp), zs.ctypes.data_as(dp), p0.ctypes.data_as(dp), C.c_double(0.01),
           iters, C.c_double(0.0), its, pout.ctypes.data_as(dp))
    assert rc == 0 and list(its) == [iters] * B
    g = glue.dll.glue_fit_run
    g.restype = C.c_int
    for b, P in enumerate(Ps):
        it, ate = C.c_int(0), C.c_double(0.0)
        trace, single, dummy = np.zeros((2, iters + 2), order="F"), np.zeros_like(p0[b]), np.zeros(1)
        Xb, w = np.asfortranarray(P["X"]), np.ones(n)
        rc = g(0, 1, n, p, P["y"].ctypes.data_as(dp), Xb.ctypes.data_as(dp), P["z"].ctypes.data_as(dp), w.ctypes.data_as(dp),
               p0[b].ctypes.data_as(dp), C.c_double(0.01), C.c_double(0.0), iters, C.c_double(0.0), C.byref(it), trace.ctypes.data_as(dp),
               single.ctypes.data_as(dp), 0, dummy.ctypes.data_as(dp), dummy.ctypes.data_as(dp), C.byref(ate))
        assert rc == 0 and cs.rel(pout[b], single, 1e-9) < 1e-9
