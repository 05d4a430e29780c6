"""Known-answer tests that pin the oracle (SURVEY.md 8c "KATs the build should create").
The reference ships no tests for this path, so these are self-consistency anchors."""
import copy
import math
import os

import numpy as np
import pytest

from oracle import causalstump_oracle as o
from tests import _cases as cs


def test_closed_form_n1():
    # n=1: K = e^lm + e^la z^2, K_n = K + e^sigma/w, LML closed form
    par = dict(sigma=0.3, lambdam=-0.2, lambdaa=0.1, Lm=np.array([-0.1]), La=np.array([0.1]), mu=0.4)
    y, X, z, w = np.array([1.7]), np.array([[0.3]]), np.array([1.0]), np.array([2.0])
    g, st, mu = o.eval_lml_grad("GP", y, X, z, w, par)
    kn = math.exp(-0.2) + math.exp(0.1) + math.exp(0.3) / 2.0
    yb = 1.7 - 0.4
    assert abs(st[1] - (-0.5 * (math.log(2 * math.pi) + math.log(kn) + yb * yb / kn))) < 1e-14
    assert abs(mu - 0.5 * 1.7) < 1e-14  # 0.5 * (y/kn) / (1/kn), quirk Q1
    assert abs(g["mu"] - yb / kn) < 1e-14
    assert abs(g["Lm"][0]) < 1e-15  # zero distance on the diagonal


def test_closed_form_n2():
    par = dict(sigma=-0.5, lambdam=0.0, lambdaa=-1.0, Lm=np.array([0.2]), La=np.array([-0.3]), mu=0.0)
    X = np.array([[0.0], [1.0]]); z = np.array([1.0, 0.0]); y = np.array([0.5, -0.25]); w = np.ones(2)
    Kl = o.kernmat_GP_SE_cpp(X, X, z, z, par)
    km01 = math.exp(0.0 - math.exp(-0.2))  # no 1/2 in the exponent (quirk Q5)
    assert abs(Kl["Km"][0, 1] - km01) < 1e-15
    assert Kl["Ka"][0, 1] == 0.0 and abs(Kl["Ka"][0, 0] - math.exp(-1.0)) < 1e-15
    inv = o.invkernel_cpp(z, w, Kl["Kmat"], par)
    Kn = Kl["Kmat"] + math.exp(-0.5) * np.eye(2)
    assert np.allclose(inv @ Kn, np.eye(2), atol=1e-14)


@pytest.mark.parametrize("n,p", [(60, 3), (90, 25)])
def test_gp_gradient_is_fd_of_lml(n, p):
    P = cs.make_problem("GP", n, p, weights=True)
    g, st, _ = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])

    def lml(par):
        return o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], par)[1][1]

    h = 1e-6
    for name in ("sigma", "lambdam", "lambdaa", "mu"):
        p1, p2 = copy.deepcopy(P["par"]), copy.deepcopy(P["par"])
        p1[name] += h; p2[name] -= h
        fd = (lml(p1) - lml(p2)) / (2 * h)
        assert abs(fd - g[name]) <= 1e-6 * max(1.0, abs(g[name])), name
    for name in ("Lm", "La"):
        for i in (0, p - 1):
            p1, p2 = copy.deepcopy(P["par"]), copy.deepcopy(P["par"])
            p1[name][i] += h; p2[name][i] -= h
            fd = (lml(p1) - lml(p2)) / (2 * h)
            assert abs(fd - g[name][i]) <= 1e-6 * max(1.0, abs(g[name][i])), (name, i)


def test_tp_gradient_follows_the_literal_formula():
    # the TP "gradient" is NOT the FD of the reported LML (extra
    # 1/exp(lambda0) in TP_adjust).  Check the literal formula instead.
    P = cs.make_problem("TP", 50, 3, nu=20.0)
    par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
    Sl = o.kernmat_TP_SE_cpp(X, X, z, z, par)
    inv = o.invkernel_cpp(z, w, Sl["Smat"], par)
    gl = o.grad_TP_SE_cpp(y, X, z, w, Sl["Smat"], Sl["Sm"], Sl["Sa"], inv, par)
    ybar = y - par["mu"]
    a = inv @ ybar
    M = float(ybar @ a)
    nu, l0 = math.exp(par["nu"]), math.exp(par["lambda0"])
    c = (nu + 50) / ((nu + M) * l0)
    assert abs(gl["gradients"]["mu"] - c * a.sum()) < 1e-12 * abs(c * a.sum())
    assert gl["gradients"]["nu"] == 0.0
    T = inv - c * np.outer(a, a)
    assert abs(gl["gradients"]["lambda0"] + 0.5 * np.sum(T * Sl["Sa"])) < 1e-10


def test_optimizer_single_steps():
    from collections import OrderedDict
    par = OrderedDict(a=1.0, b=np.array([2.0, -1.0]))
    g = OrderedDict(a=0.5, b=np.array([-0.25, 4.0]))
    z = o.list2zero(par)
    out = o.Adam_cpp(1, 0.1, 0.9, 0.999, 1e-8, z, o.list2zero(par), g, par)
    # first Adam step is lr * sign(g) (ascent), up to eps
    assert abs(out["parameters"]["a"] - 1.1) < 1e-7
    assert np.allclose(out["parameters"]["b"], [1.9, -0.9], atol=1e-7)
    out = o.Nadam_cpp(1, 0.1, 0.9, 0.999, 1e-8, z, o.list2zero(par), g, par)
    # Nadam numerator: (b1*m' + (1-b1) g)/(1-b1) with m' = (1-b1) g  ->  (b1+1) g
    assert abs(out["parameters"]["a"] - (1.0 + 0.1 * 1.9)) < 1e-6
    # Nesterov adds the OLD velocity: the first step is a no-op (quirk Q3)
    out = o.Nesterov_cpp(0.1, 0.5, z, g, par)
    assert out["parameters"]["a"] == 1.0 and abs(out["nu"]["a"] - 0.05) < 1e-16
    out2 = o.Nesterov_cpp(0.1, 0.5, out["nu"], g, out["parameters"])
    assert abs(out2["parameters"]["a"] - 1.05) < 1e-16


def test_list2zero_includes_mu_and_convergence_rule():
    P = cs.make_problem("GP", 40, 2)
    assert list(o.list2zero(P["par"]).keys()) == list(o.GP_ORDER)
    d = P["raw"]
    fit = o.CausalStump(d["y"], d["X"], d["z"], maxiter=50, tol=1e9, init_draws=(0.5, 0.7))
    assert fit.iters == 4  # |dLML| < tol only counts once iter > 3 (R/main.R:83)
    assert fit.stats.shape == (2, 6) and fit.stats[1, 0] == 0.0


def test_mu_is_overwritten_by_half_solution():
    P = cs.make_problem("GP", 40, 2)
    d = P["raw"]
    fit = o.CausalStump(d["y"], d["X"], d["z"], maxiter=1, tol=0.0, init_draws=(0.5, 0.7))
    K = o.KernelClass_GP_SE(2, np.ones(40)); K.parainit(P["y"], 0.5, 0.7)
    K.getinv_kernel(P["X"], P["z"])
    assert abs(fit.Kernel.parameters["mu"] - o.mu_solution_cpp(P["y"], P["z"], K.invKmatn, K.parameters)) < 1e-14


def test_norm_variables_quirks():
    X = np.column_stack([np.array([0., 1, 1, 0]), np.array([1., 2, 2, 1]), np.array([0.2, 0.9, 0.5, 0.1]),
                         np.array([-3., 1, 2, 4])])
    nr = o.norm_variables(np.array([1., 2, 3, 5]), X)
    m = nr["moments"]
    assert m["meanX"][0] == 0 and m["varX"][0] == 1       # binary untouched
    assert m["meanX"][1] == 1.5 and m["varX"][1] == 4.0   # {1,2}-coded counts as non-binary (Q9)
    assert m["meanX"][2] == 0 and m["varX"][2] == 1       # within [0,1] untouched
    assert m["meanX"][3] == 1.0 and m["varX"][3] == 16.0  # max-abs of the UNCENTRED column
    assert m["meanY"] == 0.0 and abs(nr["y"][0] - 1.0 / math.sqrt(np.var([1., 2, 3, 5], ddof=1))) < 1e-15


def test_predict_treat_identities():
    P = cs.make_problem("GP", 50, 3)
    K = P["K"]; K.getinv_kernel(P["X"], P["z"])
    X2 = P["X"][:17] * 0.9
    tr = K.predict_treat(P["y"], P["X"], P["z"], X2)
    Ka = K.kernel_mat(X2, P["X"], np.ones(17), P["z"])["Ka"]
    c = Ka.sum(axis=0)
    closed = K.kernel_mat(X2, X2, np.ones(17), np.ones(17))["Ka"].sum() - c @ K.invKmatn @ c
    assert abs(closed - tr["cov"].sum()) < 1e-10 * abs(closed)


def test_oracle_reproduces_golden_fixtures():
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz")
    G = np.load(path)
    names = sorted({k.split("/")[0] for k in G.files})
    assert names == ["gp_hill", "gp_nesterov", "gp_readme", "tp_hill"]
    for nm in names:
        kind = "GP" if G[nm + "/meta"][0] == 0 else "TP"
        p = int(G[nm + "/meta"][2])
        par = o.unpack_params(G[nm + "/init_params"], kind, p)
        g, st, mu = o.eval_lml_grad(kind, G[nm + "/y"], G[nm + "/X"], G[nm + "/z"], G[nm + "/w"], par)
        assert cs.rel(o.pack_params(g), G[nm + "/grad"], 1e-12) < 1e-9
        assert cs.rel(st, G[nm + "/stats"]) < 1e-11
        assert abs(mu - float(G[nm + "/mu_sol"])) < 1e-12


def test_table_exp():
    """csd::exp_tab (cs_dev.cuh) restated in NumPy: 64-entry table of 2^(j/64) + degree-5 polynomial, against 50-digit
    arithmetic on the argument range of the Gram / gradient kernels."""
    from decimal import Decimal, getcontext
    getcontext().prec = 50
    T = np.array([float(Decimal(2) ** (Decimal(j) / 64)) for j in range(64)])
    inv, hi, lo, magic = 92.33248261689366, 0.01083042469326756, 2.9815858269852933e-12, 6755399441055744.0

    def exp_tab(x):
        t = x * inv + magic
        kd = t - magic
        k = kd.astype(np.int64)
        r = (x - kd * hi) - kd * lo
        p = r / 120.0 + 1.0 / 24.0
        for c in (1.0 / 6.0, 0.5, 1.0, 1.0):
            p = p * r + c
        return np.where(x < -700.0, 0.0, np.ldexp(T[k & 63] * p, (k >> 6).astype(np.int64)))

    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-700, 5, 4000), rng.uniform(-40, 3, 4000), rng.uniform(-1e-3, 1e-3, 500), [0.0, -745.0]])
    ref = np.array([float(Decimal(float(v)).exp()) for v in x])
    got = exp_tab(x)
    ok = ref > 1e-300
    assert np.max(np.abs(got[ok] - ref[ok]) / ref[ok]) < 1e-15
    assert got[-1] == 0.0 and got[-2] == 1.0
    # the split of ln2/64 used by the kernel: kd * hi must be exact for |kd| < 2^17
    import math
    assert (math.frexp(hi)[0] * 2.0 ** 33).is_integer()
    assert abs(Decimal(hi) + Decimal(lo) - Decimal(2).ln() / 64) < Decimal(10) ** -26
