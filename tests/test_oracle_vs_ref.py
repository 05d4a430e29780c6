"""Pins oracle/causalstump_oracle.py against the reference's OWN C++ compiled here:
oracle/_ref/libcausalstump_ref.so = /root/reference/src/{kernel_GP_SE,kernel_TP_SE,optimizer}_cpp.cpp,
unmodified, built by oracle/Makefile against oracle/shim/RcppArmadillo.h (R, Rcpp and Armadillo are
not installed; the shim restates only the third-party container / LAPACK-level calls).  Where the
library cannot be built (no /root/reference, e.g. on the GPU box without the prebuilt .so) these
tests skip and tests/test_golden_ref.py covers the same ground from committed vectors."""
import numpy as np
import pytest

from oracle import causalstump_oracle as o
from oracle import ref_binding as rb
from tests import _cases as cs


@pytest.fixture(scope="module")
def ref():
    if rb.build() is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return rb.RefLib()


CASES = [("GP", 120, 1, True, False), ("GP", 97, 5, False, True), ("TP", 150, 4, False, False), ("GP", 260, 25, False, False),
         ("TP", 203, 25, False, True)]


@pytest.mark.parametrize("kind,n,p,readme,weights", CASES)
def test_every_kernel_routine_matches_the_compiled_reference(ref, kind, n, p, readme, weights):
    P = cs.make_problem(kind, n, p, readme=readme, weights=weights)
    par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
    k = 0 if kind == "GP" else 1
    flat = o.pack_params(par)
    # kernmat_*_SE_cpp, square and cross shapes
    K, Km, Ka = ref.kernmat(k, X, X, z, z, par["Lm"], par["La"], par["lambdam"], par["lambdaa" if k == 0 else "lambda0"])
    Ko = (o.kernmat_GP_SE_cpp if k == 0 else o.kernmat_TP_SE_cpp)(X, X, z, z, par)
    names = ("Kmat", "Km", "Ka") if k == 0 else ("Smat", "Sm", "Sa")
    for got, nm in zip((K, Km, Ka), names):
        assert cs.relm(Ko[nm], got) < 1e-14
    X2, z2 = X[: n // 3] * 0.9 + 0.01, 1.0 - z[: n // 3]
    Kc = ref.kernmat(k, X2, X, z2, z, par["Lm"], par["La"], par["lambdam"], par["lambdaa" if k == 0 else "lambda0"])
    Kco = (o.kernmat_GP_SE_cpp if k == 0 else o.kernmat_TP_SE_cpp)(X2, X, z2, z, par)
    assert cs.relm(Kco[names[0]], Kc[0]) < 1e-14 and Kc[0].shape == (n // 3, n)
    # invkernel_cpp (inv_sympd restated differently on the two sides: plain loops vs LAPACK)
    invK = ref.invkernel(z, w, K, par["sigma"])
    invKo = o.invkernel_cpp(z, w, Ko[names[0]], par)
    assert cs.relm(invKo, invK) < 1e-9
    assert np.array_equal(invK, invK.T)
    # mu_solution_cpp
    assert abs(ref.mu_solution(y, z, invK) - o.mu_solution_cpp(y, z, invK, par)) < 1e-12 * abs(o.mu_solution_cpp(y, z, invK, par))
    # grad_*_SE_cpp on IDENTICAL matrix inputs
    g, st = ref.grad(k, y, X, z, w, K, Km, Ka, invK, flat)
    go = (o.grad_GP_SE_cpp if k == 0 else o.grad_TP_SE_cpp)(y, X, z, w, K, Km, Ka, invK, par)
    gof = o.pack_params(go["gradients"])
    scale = np.maximum(np.abs(gof), 1e-6 * np.max(np.abs(gof)))
    assert np.max(np.abs(g - gof) / scale) < 1e-10
    assert abs(st[1] - go["stats"][1]) < 1e-11 * abs(go["stats"][1])
    assert abs(st[0] - go["stats"][0]) < 1e-9 * abs(go["stats"][0])
    # stats_*_SE
    st2 = ref.stats(k, p if not readme else 1, y, K, invK, flat)
    st2o = (o.stats_GP_SE if k == 0 else o.stats_TP_SE)(y, K, invK, par)
    assert abs(st2[1] - st2o[1]) < 1e-11 * abs(st2o[1]) and abs(st2[0] - st2o[0]) < 1e-9 * abs(st2o[0])
    # the chained evaluation the GPU path is compared with
    g3, st3, mu3, _ = ref.eval_lml_grad(k, y, X, z, w, flat)
    g3o, st3o, mu3o = o.eval_lml_grad(kind, y, X, z, w, par)
    g3of = o.pack_params(g3o)
    assert np.max(np.abs(g3 - g3of) / np.maximum(np.abs(g3of), 1e-6 * np.max(np.abs(g3of)))) < 1e-9
    assert abs(st3[1] - st3o[1]) < 1e-11 * abs(st3o[1]) and abs(mu3 - mu3o) < 1e-10 * abs(mu3o)


@pytest.mark.parametrize("kind", ["GP", "TP"])
@pytest.mark.parametrize("optim", ["Adam", "Nadam", "Nesterov"])
def test_optimizer_steps_match_the_compiled_reference(ref, kind, optim):
    p = 3
    rng = np.random.default_rng(17)
    k = 0 if kind == "GP" else 1
    P = cs.make_problem(kind, 30, p)
    para = P["par"]
    oc = {"Adam": 0, "Nadam": 1, "Nesterov": 2}[optim]
    m = o.list2zero(para)
    v = o.list2zero(para)
    fm, fv, fp = o.pack_params(m), o.pack_params(v), o.pack_params(para)
    for it in range(1, 5):
        grad = o.unpack_params(rng.normal(size=fp.size), kind, p)
        if optim == "Nesterov":
            out = o.Nesterov_cpp(0.01, 0.5, m, grad, para)
            m, para = out["nu"], out["parameters"]
        else:
            out = (o.Adam_cpp if optim == "Adam" else o.Nadam_cpp)(it, 0.01, 0.9, 0.999, 1e-8, m, v, grad, para)
            m, v, para = out["m"], out["v"], out["parameters"]
        fm, fv, fp = ref.opt_step(oc, k, p, it, 0.01, 0.5 if optim == "Nesterov" else 0.9, 0.999, 1e-8, fm, fv, o.pack_params(grad), fp)
        assert np.allclose(fp, o.pack_params(para), rtol=1e-14, atol=0)
        assert np.allclose(fm, o.pack_params(m), rtol=1e-14, atol=1e-300)
        if optim != "Nesterov":
            assert np.allclose(fv, o.pack_params(v), rtol=1e-14, atol=1e-300)


def test_reference_inv_sympd_rejects_non_pd(ref):
    A = np.array([[1.0, 2.0], [2.0, 1.0]])
    with pytest.raises(np.linalg.LinAlgError):
        ref.invkernel(np.zeros(2), np.ones(2), A, -40.0)
    with pytest.raises(o.NotPositiveDefinite):
        o.invkernel_cpp(np.zeros(2), np.ones(2), A, dict(sigma=-40.0))


def test_host_lapack_hook_matches_the_plain_loops(ref):
    """inv_sympd / log_det through the host LAPACK (dpotrf + dpotri, dgetrf: the calls Armadillo makes; used by the
    timing legs of bench.py) give the same routine outputs as the shim's plain loops, and still reject non-PD input."""
    P = cs.make_problem("GP", 211, 7)
    par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
    flat = o.pack_params(par)
    try:
        g0, st0, mu0, m0 = ref.eval_lml_grad(0, y, X, z, w, flat)
        ld0 = ref.log_det(m0["invK"])
        assert "openblas" in ref.use_host_lapack().lower()
        g1, st1, mu1, m1 = ref.eval_lml_grad(0, y, X, z, w, flat)
        ld1 = ref.log_det(m1["invK"])
        assert np.allclose(m1["invK"], m0["invK"], rtol=1e-9, atol=1e-11 * np.abs(m0["invK"]).max())
        assert np.allclose(g1, g0, rtol=1e-8, atol=1e-10) and np.allclose(st1, st0, rtol=1e-10)
        assert abs(mu1 - mu0) <= 1e-9 * max(1.0, abs(mu0)) and ld1[1] == ld0[1] and abs(ld1[0] - ld0[0]) <= 1e-9 * abs(ld0[0])
        with pytest.raises(np.linalg.LinAlgError):
            ref.invkernel(np.zeros(2), np.ones(2), np.array([[1.0, 2.0], [2.0, 1.0]]), -40.0)
    finally:
        ref.use_host_lapack(False)
