"""Shared parity bodies: the same checks run against the SIMT-emulated kernels (CPU,
small n) and against the CUDA library on a B200 (-m gpu, reference sizes)."""
import numpy as np

from causalstump_b200 import _lib, hill
from oracle import causalstump_oracle as o

# tolerances of BASELINE.json north_star
TOL_LML = 1e-9
TOL_GRAD = 1e-8
TOL_MAP = 1e-8


def rel(a, b, floor=1e-300):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def relm(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def make_problem(kind, n, p, replicate=0, readme=False, weights=False, nu=200.0, draws=None):
    d = hill.readme_example(n) if readme else hill.hill_surface_b(n=n, p=p, replicate=replicate)
    nr = o.norm_variables(d["y"], d["X"])
    y, X, z = nr["y"], np.asfortranarray(nr["X"]), d["z"]
    p = X.shape[1]
    w = np.random.default_rng(99).uniform(0.5, 2.0, n) if weights else np.ones(n)
    if kind == "GP":
        K = o.KernelClass_GP_SE(p, w)
        K.parainit(y, *(draws or (0.5, 0.7)))
    else:
        K = o.KernelClass_TP_SE(p, w)
        K.parainit(y, nu, *(draws or (0.5, 0.3)))
    return dict(kind=kind, raw=d, y=y, X=X, z=z, w=w, K=K, par=K.parameters, moments=nr["moments"], n=n, p=p)


def kcode(kind):
    return _lib.CS_GP if kind == "GP" else _lib.CS_TP


def check_eval(lib, kind, n, p, tile=0, **kw):
    P = make_problem(kind, n, p, **kw)
    g, st, mu = o.eval_lml_grad(kind, P["y"], P["X"], P["z"], P["w"], P["par"])
    f = _lib.Fit(lib, kcode(kind), P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=tile)
    try:
        g2, st2, mu2 = f.eval()
    finally:
        f.close()
    gref = o.pack_params(g)
    # gradients: relative to the gradient scale (entries that are ~0 by cancellation are
    # compared against the largest entry)
    gerr = float(np.max(np.abs(g2 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))))
    assert abs(st2[1] - st[1]) <= TOL_LML * abs(st[1]), ("LML", st2[1], st[1])
    assert gerr <= TOL_GRAD, ("grad", gerr)
    assert abs(st2[0] - st[0]) <= 1e-7 * abs(st[0]), ("sqres", st2[0], st[0])
    assert abs(mu2 - mu) <= TOL_GRAD * abs(mu), ("mu", mu2, mu)
    return gerr


OPT = {"Adam": _lib.CS_OPT_ADAM, "Nadam": _lib.CS_OPT_NADAM, "Nesterov": _lib.CS_OPT_NESTEROV, "GD": _lib.CS_OPT_NESTEROV}


def oracle_fit(P, optim, maxiter, tol, lr, momentum=0.0):
    d = P["raw"]
    draws = (np.exp(P["par"]["lambdam"]), np.exp(P["par"]["lambdaa" if P["kind"] == "GP" else "lambda0"]))
    return o.CausalStump(d["y"], d["X"], d["z"], w=P["w"], myoptim=optim, maxiter=maxiter, tol=tol,
                         prior=(P["kind"] == "TP"), nu=float(np.exp(P["par"].get("nu", 0.0))), learning_rate=lr,
                         momentum=momentum, init_draws=draws, faithful=False)


def check_fit(lib, kind, n, p, optim="Nadam", maxiter=10, tol=1e-12, lr=0.01, momentum=0.5, tile=0, chunk=0,
              n2=40, **kw):
    """Same number of optimizer steps on both sides, then parameters, trace, predict,
    treatment and PEHE (6 significant digits)."""
    P = make_problem(kind, n, p, **kw)
    fit = oracle_fit(P, optim, maxiter, tol, lr, momentum)
    f = _lib.Fit(lib, kcode(kind), P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), optim=OPT[optim], lr=lr,
                 momentum=(0.0 if optim == "GD" else momentum), tile=tile, chunk=chunk)
    try:
        it, trace = f.run(maxiter, tol)
        assert it == fit.iters, (it, fit.iters)
        pref = o.pack_params(fit.Kernel.parameters)
        assert rel(f.get_params(), pref, 1e-9) <= 1e-7, rel(f.get_params(), pref, 1e-9)
        assert rel(trace[1, 1:], fit.stats[1, 1:]) <= 1e-8
        assert rel(trace[0, 1:], fit.stats[0, 1:]) <= 1e-6
        # predict / treatment on new points and on the training points
        d2 = hill.readme_example(n2, seed=77) if P["raw"]["X"].shape[1] == 1 else hill.hill_surface_b(n=n2, p=P["p"], replicate=50)
        X2 = o.norm_variables(X=d2["X"], moments=fit.moments)["X"]
        z2 = d2["z"]
        y, X, z = P["y"], P["X"], P["z"]
        pr = fit.Kernel.predict(y, X, z, X2, z2)
        tr = fit.Kernel.predict_treat(y, X, z, X2)
        jit = 1e-6 if kind == "TP" else 0.0
        p2 = f.predict(X2, z2, want_cov=True)
        t2 = f.treat(X2, cov_jitter=jit, want_cov=True)
        assert rel(p2["map"], pr["map"], 1e-6) <= TOL_MAP * 10
        assert relm(p2["map"], pr["map"]) <= TOL_MAP
        assert relm(p2["var"], np.diag(pr["cov"])) <= 1e-7
        assert relm(p2["cov"], pr["cov"]) <= 1e-7
        assert relm(t2["map"], tr["map"]) <= TOL_MAP
        assert relm(t2["var"], np.diag(tr["cov"])) <= 1e-7
        assert relm(t2["cov"], tr["cov"]) <= 1e-7
        assert abs(t2["ate"] - tr["ate_map"]) <= TOL_MAP * max(abs(tr["ate_map"]), 1e-3)
        assert abs(t2["ate_var"] - np.sum(tr["cov"]) / n ** 2) <= 1e-7 * abs(np.sum(tr["cov"]) / n ** 2)
        # PEHE on the training sample, de-normalised like R/treatment.R:28 (6 significant digits)
        ttr_o = fit.Kernel.predict_treat(y, X, z, X)["map"] * np.sqrt(fit.moments["varY"])
        ttr_c = f.treat(X)["map"] * np.sqrt(fit.moments["varY"])
        pehe_o = float(np.sqrt(np.mean((P["raw"]["tau"] - ttr_o) ** 2)))
        pehe_c = float(np.sqrt(np.mean((P["raw"]["tau"] - ttr_c) ** 2)))
        assert abs(pehe_c - pehe_o) <= 0.5e-6 * abs(pehe_o), (pehe_c, pehe_o)
    finally:
        f.close()
    return pehe_c


class _raises:
    """minimal pytest.raises (this module is also imported by __graft_entry__.smoke without pytest)"""

    def __init__(self, exc):
        self.exc, self.value = exc, None

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        assert et is not None and issubclass(et, self.exc), ("expected", self.exc, "got", et)
        self.value = ev
        return True


def check_limits(emulib):
    """p = CS_MAX_P works, p > CS_MAX_P is refused; a non positive definite member of a batch is reported with its
    pivot (what makes arma::inv_sympd throw at src/kernel_GP_SE_cpp.cpp:56) instead of producing numbers."""
    # p = 40: element-wise kernels in one launch; p = 70 and p = CS_MAX_P: the gradient kernel runs one launch per window
    # of dimensions (its per-thread accumulators live in shared memory); p > CS_MAX_P is refused
    for p_ in (40, 70, _lib.CS_MAX_P):
        P = make_problem("GP", 70, p_)
        g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
        f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]))
        g2, st2, _ = f.eval()
        f.close()
        gref = o.pack_params(g)
        assert float(np.max(np.abs(g2 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref))))) < TOL_GRAD, p_
        assert abs(st2[1] - st[1]) < TOL_LML * abs(st[1])
    pbig = _lib.CS_MAX_P + 1
    with _raises(_lib.CsError) as ei:
        _lib.Fit(emulib, 0, np.zeros(10), np.zeros((10, pbig)), np.zeros(10), None, np.zeros(2 * pbig + 4))
    assert ei.value.code == -5
    # member 1: two identical rows and (numerically) no noise -> singular Gram matrix
    good = make_problem("GP", 40, 2)
    bad = make_problem("GP", 40, 2, replicate=1)
    Xb = bad["X"].copy(); Xb[7] = Xb[3]
    zb = bad["z"].copy(); zb[7] = zb[3]
    pb = o.pack_params(bad["par"]); pb[0] = -80.0
    fb = _lib.FitBatch(emulib, 0, [(good["y"], good["X"], good["z"], None, o.pack_params(good["par"])), (bad["y"], Xb, zb, None, pb)])
    with _raises(_lib.NotPositiveDefinite) as ei:
        fb.run(5, 1e-12)
    assert 1 <= ei.value.code <= 40 and "member 1" in str(ei.value)
    fb.close()
    with _raises(o.NotPositiveDefinite):
        o.eval_lml_grad("GP", bad["y"], Xb, zb, np.ones(40), o.unpack_params(pb, "GP", 2))


def check_failing_member(emulib):
    """a non positive definite member (or one whose LML turns NaN) must not discard the healthy members of
    the batch.  In the reference only the failing CausalStump() call errors (inv_sympd throws inside that call,
    src/kernel_GP_SE_cpp.cpp:56; `if (change < tol ...)` on NA, R/main.R:82)."""
    good = make_problem("GP", 40, 2)
    bad = make_problem("GP", 40, 2, replicate=1)
    nanp = make_problem("GP", 40, 2, replicate=2)
    Xb = bad["X"].copy(); Xb[7] = Xb[3]
    zb = bad["z"].copy(); zb[7] = zb[3]
    pb = o.pack_params(bad["par"]); pb[0] = -80.0
    yn = nanp["y"].copy(); yn[5] = np.nan
    single = _lib.Fit(emulib, 0, good["y"], good["X"], good["z"], None, o.pack_params(good["par"]))
    it0, tr0 = single.run(6, 1e-12)
    p0 = single.get_params()
    single.close()
    fb = _lib.FitBatch(emulib, 0, [(bad["y"], Xb, zb, None, pb), (good["y"], good["X"], good["z"], None, o.pack_params(good["par"])),
                                   (yn, nanp["X"], nanp["z"], None, o.pack_params(nanp["par"]))])
    res = fb.run(6, 1e-12, strict=False)
    st = fb.member_status()
    assert 1 <= st[0] <= 40 and st[1] == 0 and st[2] == _lib.CS_ERR_NONFINITE
    assert res[1][0] == it0 and np.array_equal(res[1][1], tr0) and np.array_equal(fb.select(1).get_params(), p0)
    assert res[0][0] <= 2 and res[2][0] == 1     # stopped at the iteration the reference would error in, not at maxiter
    with _raises(_lib.NotPositiveDefinite) as ei:
        fb.run(6, 1e-12)                         # strict: the first failing member is the error, after all members ran
    assert "member 0" in str(ei.value)
    fb.close()
    f = _lib.Fit(emulib, 0, yn, nanp["X"], nanp["z"], None, o.pack_params(nanp["par"]))
    with _raises(_lib.CsError) as ei:
        f.run(6, 1e-12)
    assert ei.value.code == _lib.CS_ERR_NONFINITE
    f.close()
