"""CPU oracle for the CausalStump GP / Student-t-process fitting hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  The product package ``causalstump_b200`` never imports it and has no
CPU fallback.

PARITY PINNED against the reference's own C++.  The reference (mazphilip/CausalStump, R + RcppArmadillo)
ships no unit tests, golden vectors or known-answer fixtures for this path (SURVEY.md
section 4 / 8c) and its R build cannot run in this image (no R, Rcpp, Armadillo), but
its three hot-path source files compile unmodified against the header shim of
oracle/shim (oracle/Makefile -> oracle/_ref/libcausalstump_ref.so).  This restatement is
held to that library routine by routine (tests/test_oracle_vs_ref.py) and to golden
vectors it produced: tests/golden/golden_ref_v1.npz (five small cases incl. optimizer
trajectories, predict / treatment) and tests/golden/golden_ref_8192.npz (n=8192 GP and
Student-t, n=4000: tests/test_golden_ref.py checks the n=4000 case here).  What stays
unpinned: the R-level glue (restated with line numbers, not executed), R's RNG stream and
mvnfast's sitmo stream (sampling parity is distributional).  Self-consistency known-answer
tests (closed forms for n=1,2, finite-difference gradients, optimizer single-step vectors)
are in tests/test_oracle.py.

Third-party arithmetic that lives outside /root/reference and is restated here
(versions unpinned by the reference, DESCRIPTION:10-15):
  * RcppArmadillo ``inv_sympd``  -> LAPACK dpotrf('L') + dpotri('L') + mirror
  * RcppArmadillo ``log_det``    -> LAPACK dgetrf, sum(log|u_ii|) and sign
  * RcppArmadillo ``trace(A*B)`` -> sum_ij A_ij B_ji (O(n^2))
  * mvnfast ``rmvt`` (>= 0.2.0)  -> z @ chol(cov)^T / sqrt(chi2_df/df); its
    sitmo RNG stream is not reproducible, so sampling parity is distributional

Every function cites the reference file:line it follows.  All arrays are fp64.
Parameter containers are ordered dicts in the reference list order
(GP: sigma, lambdam, lambdaa, Lm, La, mu -- R/kernel_GP_SE_class.R:13-19;
 TP: sigma, lambdam, nu, lambda0, Lm, La, mu -- R/kernel_TP_SE_class.R:14-21).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
from scipy.linalg import lapack
from scipy.special import gammaln

GP_ORDER = ("sigma", "lambdam", "lambdaa", "Lm", "La", "mu")
TP_ORDER = ("sigma", "lambdam", "nu", "lambda0", "Lm", "La", "mu")


class NotPositiveDefinite(RuntimeError):
    """arma::inv_sympd failure (src/kernel_GP_SE_cpp.cpp:56)."""

    def __init__(self, pivot):
        super().__init__(f"inv_sympd(): matrix is singular or not positive definite (pivot {pivot})")
        self.pivot = int(pivot)


# --------------------------------------------------------------------------
# R/utilities.R
# --------------------------------------------------------------------------
def r_var(x):
    """stats::var -- unbiased (n-1) sample variance."""
    x = np.asarray(x, dtype=np.float64)
    return float(np.var(x, ddof=1))


def norm_variables(y=None, X=None, moments=None):
    """R/utilities.R:1-33.  Returns dict(moments=, y=, X=) or dict(X=)."""
    X = np.array(X, dtype=np.float64, copy=True)
    if X.ndim == 1:
        X = X[:, None]
    p = X.shape[1]
    if moments is None:
        moments = dict(meanX=np.zeros(p), varX=np.ones(p), meanY=0.0, varY=r_var(y))  # :6-9
        for j in range(p):  # :12-17
            col = X[:, j]
            col = col[~np.isnan(col)]
            binary = bool(np.all((col == 0.0) | (col == 1.0)))
            within_unit = bool(np.all((col <= 1.0) & (col >= 0.0)))
            if (not binary) and (not within_unit):
                moments["meanX"][j] = col.mean()  # :21
                moments["varX"][j] = np.max(np.abs(X[:, j])) ** 2  # :22 (uncentred max-abs)
    Xn = (X - moments["meanX"][None, :]) / np.sqrt(moments["varX"])[None, :]  # :24-25
    if y is None:
        return dict(X=Xn)
    yn = (np.asarray(y, dtype=np.float64) - moments["meanY"]) / math.sqrt(moments["varY"])  # :29
    return dict(moments=moments, y=yn, X=Xn)


# --------------------------------------------------------------------------
# src/kernel_GP_SE_cpp.cpp / src/kernel_TP_SE_cpp.cpp
# --------------------------------------------------------------------------
def _kernmat(X1, X2, Z1, Z2, Lm, La, lam_m, lam_a):
    """Shared body of kernmat_GP_SE_cpp (:72-140) and kernmat_TP_SE_cpp (:46-90)."""
    X1 = np.asarray(X1, dtype=np.float64)
    X2 = np.asarray(X2, dtype=np.float64)
    if X1.ndim == 1:
        X1 = X1[:, None]
    if X2.ndim == 1:
        X2 = X2[:, None]
    n1, n2, p = X1.shape[0], X2.shape[0], X2.shape[1]
    Km = np.zeros((n1, n2))
    Ka = np.zeros((n1, n2))
    Lm = np.atleast_1d(np.asarray(Lm, dtype=np.float64))
    La = np.atleast_1d(np.asarray(La, dtype=np.float64))
    for i in range(p):  # :105-119, one pass per dimension
        d2 = (X1[:, i][:, None] - X2[:, i][None, :]) ** 2
        Km += d2 * math.exp(-Lm[i])
        Ka += d2 * math.exp(-La[i])
    Z1 = np.asarray(Z1, dtype=np.float64)
    Z2 = np.asarray(Z2, dtype=np.float64)
    Km = np.exp(lam_m - Km)  # :130
    Ka = np.exp(lam_a - Ka) * Z1[:, None] * Z2[None, :]  # :131
    return Km + Ka, Km, Ka  # :132


def kernmat_GP_SE_cpp(X1, X2, Z1, Z2, para):
    """src/kernel_GP_SE_cpp.cpp:72-140 -> dict(Kmat, Km, Ka)."""
    K, Km, Ka = _kernmat(X1, X2, Z1, Z2, para["Lm"], para["La"], para["lambdam"], para["lambdaa"])
    return OrderedDict(Kmat=K, Km=Km, Ka=Ka)


def kernmat_TP_SE_cpp(X1, X2, Z1, Z2, para):
    """src/kernel_TP_SE_cpp.cpp:46-90 -> dict(Smat, Sm, Sa); lambda0 replaces lambdaa."""
    K, Km, Ka = _kernmat(X1, X2, Z1, Z2, para["Lm"], para["La"], para["lambdam"], para["lambda0"])
    return OrderedDict(Smat=K, Sm=Km, Sa=Ka)


def inv_sympd(A):
    """arma::inv_sympd restated: dpotrf('L') + dpotri('L') + mirror lower->upper."""
    c, info = lapack.dpotrf(np.asfortranarray(A), lower=1, clean=0, overwrite_a=0)
    if info != 0:
        raise NotPositiveDefinite(info)
    inv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    if info != 0:
        raise NotPositiveDefinite(info)
    inv = np.tril(inv) + np.tril(inv, -1).T
    return inv


def log_det(A):
    """arma::log_det restated: dgetrf, val = sum(log|u_ii|), sign from pivots/diag."""
    lu, piv, info = lapack.dgetrf(np.asfortranarray(A), overwrite_a=0)
    d = np.diag(lu)
    val = float(np.sum(np.log(np.abs(d))))
    sign = float(np.prod(np.sign(d))) * (-1.0) ** int(np.sum(piv != np.arange(len(piv))))
    return val, sign


def invkernel_cpp(z, w, mymat, parameters):
    """src/kernel_GP_SE_cpp.cpp:46-59: diag += exp(sigma)/w ; inv_sympd."""
    A = np.array(mymat, dtype=np.float64, copy=True)
    w = np.asarray(w, dtype=np.float64)
    A[np.diag_indices_from(A)] += math.exp(parameters["sigma"]) / w  # :55
    return inv_sympd(A)  # :56


def mu_solution_cpp(y, z, invKmat, parameters):
    """src/kernel_GP_SE_cpp.cpp:61-70 (note the 0.5, quirk Q1)."""
    y = np.asarray(y, dtype=np.float64)
    return float(0.5 * np.sum(invKmat @ y) / np.sum(np.sum(invKmat, axis=0)))  # :64


def _trace_prod(A, B):
    """trace(A*B) as Armadillo evaluates it: sum_ij A_ij B_ji."""
    return float(np.sum(A * B.T))


def evid_grad(mymat, dK):
    """src/kernel_GP_SE_cpp.cpp:8-14."""
    return -0.5 * _trace_prod(mymat, dK)


def evid_scale_gradients(X, tmpK, Km, Ka, parameters):
    """src/kernel_GP_SE_cpp.cpp:16-44 (and the TP twin, kernel_TP_SE_cpp.cpp:15-43)."""
    n, p = X.shape
    Lm = np.atleast_1d(np.asarray(parameters["Lm"], dtype=np.float64))
    La = np.atleast_1d(np.asarray(parameters["La"], dtype=np.float64))
    grad_m = np.zeros(p)
    grad_a = np.zeros(p)
    for i in range(p):  # :28-41
        tmpX = (X[:, i][None, :] - X[:, i][:, None]) ** 2  # :30-32
        tmpM = Km * tmpX * math.exp(-Lm[i])  # :35
        tmpA = Ka * tmpX * math.exp(-La[i])  # :36
        grad_m[i] = evid_grad(tmpK, tmpM)
        grad_a[i] = evid_grad(tmpK, tmpA)
    return grad_m, grad_a


def _as_matrix(X):
    X = np.asarray(X, dtype=np.float64)
    return X[:, None] if X.ndim == 1 else X


def grad_GP_SE_cpp(y, X, z, w, Kmat, Km, Ka, invKmatn, parameters):
    """src/kernel_GP_SE_cpp.cpp:143-202 -> dict(gradients=, stats=[sqres, LML])."""
    y = np.asarray(y, dtype=np.float64)
    X = _as_matrix(X)
    w = np.asarray(w, dtype=np.float64)
    n = X.shape[0]
    gradients = OrderedDict((k, None) for k in parameters.keys())  # clone(parameters) :145
    ybar = y - parameters["mu"]  # :157
    alpha = invKmatn @ ybar  # :158
    tmpK = invKmatn - np.outer(alpha, alpha)  # :159
    dK = np.diag(math.exp(parameters["sigma"]) / w)  # :162
    gradients["sigma"] = evid_grad(tmpK, dK)  # :163
    gradients["lambdam"] = evid_grad(tmpK, Km)  # :170
    gradients["lambdaa"] = evid_grad(tmpK, Ka)  # :173
    gm, ga = evid_scale_gradients(X, tmpK, Km, Ka, parameters)  # :176
    gradients["Lm"] = gm
    gradients["La"] = ga
    gradients["mu"] = float(np.sum(invKmatn @ ybar))  # :184
    stats = np.zeros(2)
    stats[0] = float(np.linalg.norm(y - (Kmat @ alpha + parameters["mu"])) ** 2)  # :192
    val, _sign = log_det(invKmatn)  # :197 logdet of the INVERSE
    stats[1] = -0.5 * (n * math.log(2.0 * math.pi) - val + float(ybar @ alpha))  # :198
    return dict(gradients=gradients, stats=stats)


def stats_GP_SE(y, Kmat, invKmatn, parameters):
    """src/kernel_GP_SE_cpp.cpp:204-227."""
    y = np.asarray(y, dtype=np.float64)
    n = y.size
    ybar = y - parameters["mu"]
    alpha = invKmatn @ ybar
    stats = np.zeros(2)
    stats[0] = float(np.linalg.norm(y - (Kmat @ alpha + parameters["mu"])) ** 2)  # :218
    val, _sign = log_det(invKmatn)
    stats[1] = -0.5 * (n * math.log(2.0 * math.pi) - val + float(ybar @ alpha))  # :224
    return stats


def grad_TP_SE_cpp(y, X, z, w, Smat, Sm, Sa, invSmatn, parameters):
    """src/kernel_TP_SE_cpp.cpp:100-163 (TP_adjust carries 1/exp(lambda0), quirk Q2)."""
    y = np.asarray(y, dtype=np.float64)
    X = _as_matrix(X)
    w = np.asarray(w, dtype=np.float64)
    n = X.shape[0]
    gradients = OrderedDict((k, None) for k in parameters.keys())
    lambda0 = math.exp(parameters["lambda0"])  # :116
    nu = math.exp(parameters["nu"])  # :117
    ybar = y - parameters["mu"]
    alpha = invSmatn @ ybar  # :120
    M_norm = float(ybar @ alpha)  # :121
    TP_adjust = (nu + n) / ((nu + M_norm) * lambda0)  # :122
    tmpS = invSmatn - TP_adjust * np.outer(alpha, alpha)  # :124
    gradients["mu"] = TP_adjust * float(np.sum(invSmatn @ ybar))  # :127
    gradients["nu"] = 0.0  # :130
    gradients["lambda0"] = evid_grad(tmpS, Sa)  # :133
    dS = np.diag(math.exp(parameters["sigma"]) / w)  # :137
    gradients["sigma"] = evid_grad(tmpS, dS)  # :138
    gradients["lambdam"] = evid_grad(tmpS, Sm)  # :145
    gm, ga = evid_scale_gradients(X, tmpS, Sm, Sa, parameters)  # :148
    gradients["Lm"] = gm
    gradients["La"] = ga
    stats = np.zeros(2)
    stats[0] = float(np.linalg.norm(y - (Smat @ alpha + parameters["mu"])) ** 2)  # :153
    val, _sign = log_det(invSmatn)  # :158
    stats[1] = float(
        gammaln(0.5 * (nu + n)) - gammaln(0.5 * nu)
        - 0.5 * (n * math.log(nu * math.pi) - val + (nu + n) * math.log(1.0 + M_norm / nu))
    )  # :159
    return dict(gradients=gradients, stats=stats)


def stats_TP_SE(y, Kmat, invKmatn, parameters):
    """src/kernel_TP_SE_cpp.cpp:165-192."""
    y = np.asarray(y, dtype=np.float64)
    n = y.size
    ybar = y - parameters["mu"]
    alpha = invKmatn @ ybar
    nu = math.exp(parameters["nu"])
    M_norm = float(ybar @ alpha)
    stats = np.zeros(2)
    stats[0] = float(np.linalg.norm(y - (Kmat @ alpha + parameters["mu"])) ** 2)  # :183
    val, _sign = log_det(invKmatn)
    stats[1] = float(
        gammaln(0.5 * (nu + n)) - gammaln(0.5 * nu)
        - 0.5 * (n * math.log(nu * math.pi) - val + (nu + n) * math.log(1.0 + M_norm / nu))
    )  # :189
    return stats


# --------------------------------------------------------------------------
# src/optimizer_cpp.cpp -- positional loops over the first len(state) entries
# --------------------------------------------------------------------------
def _vals(d):
    return [np.array(v, dtype=np.float64, copy=True) for v in d.values()]


def _rebuild(template, vals):
    out = OrderedDict()
    for (k, old), v in zip(template.items(), vals):
        out[k] = float(v) if np.ndim(old) == 0 else np.asarray(v, dtype=np.float64)
    return out


def Nesterov_cpp(learn_rate, momentum, nu, grad, para):
    """src/optimizer_cpp.cpp:9-26.  Adds the OLD velocity (quirk Q3)."""
    nu_v, g_v, p_v = _vals(nu), _vals(grad), _vals(para)
    out_nu, out_p = list(nu_v), list(p_v)
    for i in range(len(nu_v)):
        out_nu[i] = momentum * nu_v[i] + learn_rate * g_v[i]  # :19
        out_p[i] = p_v[i] + nu_v[i]  # :20
    return OrderedDict(nu=_rebuild(nu, out_nu), parameters=_rebuild(para, out_p))


def Nadam_cpp(it, learn_rate, beta1, beta2, eps, m, v, grad, para):
    """src/optimizer_cpp.cpp:28-44 (ascent; quirk Q4)."""
    m_v, v_v, g_v, p_v = _vals(m), _vals(v), _vals(grad), _vals(para)
    om, ov, op = list(m_v), list(v_v), list(p_v)
    it = float(it)
    for i in range(len(m_v)):
        om[i] = beta1 * m_v[i] + (1 - beta1) * g_v[i]  # :39
        ov[i] = beta2 * v_v[i] + (1 - beta2) * g_v[i] ** 2  # :40
        op[i] = p_v[i] + learn_rate * ((beta1 * om[i] + (1 - beta1) * g_v[i]) / (1 - beta1 ** it)) / (
            np.sqrt(ov[i] / (1 - beta2 ** it)) + eps
        )  # :41
    return OrderedDict(m=_rebuild(m, om), v=_rebuild(v, ov), parameters=_rebuild(para, op))


def Adam_cpp(it, learn_rate, beta1, beta2, eps, m, v, grad, para):
    """src/optimizer_cpp.cpp:46-63."""
    m_v, v_v, g_v, p_v = _vals(m), _vals(v), _vals(grad), _vals(para)
    om, ov, op = list(m_v), list(v_v), list(p_v)
    it = float(it)
    for i in range(len(m_v)):
        om[i] = beta1 * m_v[i] + (1 - beta1) * g_v[i]  # :57
        ov[i] = beta2 * v_v[i] + (1 - beta2) * g_v[i] ** 2  # :58
        op[i] = p_v[i] + learn_rate * (om[i] / (1 - beta1 ** it)) / (np.sqrt(ov[i] / (1 - beta2 ** it)) + eps)  # :59
    return OrderedDict(m=_rebuild(m, om), v=_rebuild(v, ov), parameters=_rebuild(para, op))


# --------------------------------------------------------------------------
# R/optimizer_classes.R
# --------------------------------------------------------------------------
def list2zero(mylist):
    """R/optimizer_classes.R:1-11 -- zeros shaped like ``parameters`` (incl. mu)."""
    return OrderedDict((k, 0.0 if np.ndim(v) == 0 else np.zeros_like(np.asarray(v, dtype=np.float64)))
                       for k, v in mylist.items())


class optAdam:
    """R/optimizer_classes.R:14-34 (eps hard-coded 1e-8, :23)."""

    def __init__(self, lr, beta1, beta2):
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2
        self.m = self.v = None

    def initOpt(self, kernel):
        self.m = list2zero(kernel.parameters)
        self.v = list2zero(kernel.parameters)

    def update(self, it, parameters, gradients):
        out = Adam_cpp(it, self.lr, self.beta1, self.beta2, 1e-8, self.m, self.v, gradients, parameters)
        self.m, self.v = out["m"], out["v"]
        return out["parameters"]


class optNadam(optAdam):
    """R/optimizer_classes.R:36-53."""

    def update(self, it, parameters, gradients):
        out = Nadam_cpp(it, self.lr, self.beta1, self.beta2, 1e-8, self.m, self.v, gradients, parameters)
        self.m, self.v = out["m"], out["v"]
        return out["parameters"]


class optNestorov:
    """R/optimizer_classes.R:55-68."""

    def __init__(self, lr, momentum):
        self.lr, self.momentum = lr, momentum
        self.nu = None

    def initOpt(self, kernel):
        self.nu = list2zero(kernel.parameters)

    def update(self, it, parameters, gradients):
        out = Nesterov_cpp(self.lr, self.momentum, self.nu, gradients, parameters)
        self.nu = out["nu"]
        return out["parameters"]


# --------------------------------------------------------------------------
# R/kernel_GP_SE_class.R / R/kernel_TP_SE_class.R
# --------------------------------------------------------------------------
class KernelClass_GP_SE:
    """R/kernel_GP_SE_class.R:1-107."""

    kind = "GP"

    def __init__(self, p, w):
        self.p = int(p)
        self.w = np.asarray(w, dtype=np.float64)
        self.parameters = None
        self.invKmatn = self.Kmat = self.Km = self.Ka = None

    def parainit(self, y, lambdam_init, lambdaa_init):
        """:10-22.  The two runif(1,.2,1) draws are inputs (R's RNG is unavailable)."""
        self.parameters = OrderedDict(
            sigma=math.log(r_var(y)),
            lambdam=math.log(lambdam_init),
            lambdaa=math.log(lambdaa_init),
            Lm=np.full(self.p, -0.1),
            La=np.full(self.p, 0.1),
            mu=float(np.mean(y)),
        )

    def kernel_mat(self, X1, X2, z1, z2):  # :23-30
        return kernmat_GP_SE_cpp(_as_matrix(X1), _as_matrix(X2), z1, z2, self.parameters)

    def getinv_kernel(self, X, z):  # :31-44
        Kl = self.kernel_mat(X, X, z, z)
        self.Kmat, self.Km, self.Ka = Kl["Kmat"], Kl["Km"], Kl["Ka"]
        self.invKmatn = invkernel_cpp(z, self.w, self.Kmat, self.parameters)

    def para_update(self, it, y, X, z, Optim, faithful=True):  # :45-60
        self.getinv_kernel(X, z)
        gl = grad_GP_SE_cpp(y, _as_matrix(X), z, self.w, self.Kmat, self.Km, self.Ka, self.invKmatn, self.parameters)
        self.parameters = Optim.update(it, self.parameters, gl["gradients"])
        self.mean_solution(y, z)
        if faithful:
            self.get_train_stats(y, X, z)  # :58, result discarded
        return gl["stats"]

    def get_train_stats(self, y, X, z):  # :61-67
        self.getinv_kernel(X, z)
        return stats_GP_SE(y, self.Kmat, self.invKmatn, self.parameters)

    def mean_solution(self, y, z):  # :68-73 (uses the pre-update inverse, quirk Q1)
        self.parameters["mu"] = mu_solution_cpp(y, z, self.invKmatn, self.parameters)

    def predict(self, y, X, z, X2, z2):  # :74-85
        X2 = _as_matrix(X2)
        n2 = X2.shape[0]
        K_xX = self.kernel_mat(X2, X, z2, z)["Kmat"]
        K_xx = self.kernel_mat(X2, X2, z2, z2)["Kmat"]
        mu = self.parameters["mu"]
        map_ = K_xX @ self.invKmatn @ (np.asarray(y) - mu) + mu  # :81
        cov = K_xx - K_xX @ self.invKmatn @ K_xX.T + math.exp(self.parameters["sigma"]) * np.eye(n2)  # :82
        ci = np.column_stack([-1.96 * np.diag(cov), 1.96 * np.diag(cov)])  # :83 (variance, quirk Q8)
        return dict(map=map_, ci=ci, cov=cov)

    def predict_treat(self, y, X, z, X2):  # :86-105
        n = len(y)
        X2 = _as_matrix(X2)
        n2 = X2.shape[0]
        z2 = np.ones(n2)
        K_xX = self.kernel_mat(X2, X, z2, z)["Ka"]
        K_xx = self.kernel_mat(X2, X2, z2, z2)["Ka"]
        map_ = K_xX @ self.invKmatn @ (np.asarray(y) - self.parameters["mu"])  # :95
        ate = float(np.mean(map_))
        cov = K_xx - K_xX @ self.invKmatn @ K_xX.T  # :99
        ci = np.column_stack([-1.96 * np.diag(cov), 1.96 * np.diag(cov)])
        ate_cov = float(np.sum(cov)) / (n ** 2)  # :101 (training n, quirk Q8)
        ate_ci = np.array([-1.96 * math.sqrt(ate_cov), 1.96 * math.sqrt(ate_cov)])
        return dict(map=map_, ci=ci, ate_map=ate, ate_ci=ate_ci, cov=cov)


def rmvt(n, sigma, df, rng):
    """mvnfast::rmvt(n, mu=0, sigma, df) restated: N(0,sigma) / sqrt(chi2_df/df)."""
    L = np.linalg.cholesky(sigma)
    zn = rng.standard_normal((n, sigma.shape[0])) @ L.T
    g = rng.chisquare(df, size=n) / df
    return zn / np.sqrt(g)[:, None]


class KernelClass_TP_SE:
    """R/kernel_TP_SE_class.R:1-127."""

    kind = "TP"

    def __init__(self, p, w, nsampling=5000):
        self.p = int(p)
        self.w = np.asarray(w, dtype=np.float64)
        self.nsampling = nsampling
        self.parameters = None
        self.invSmatn = self.Smat = self.Sm = self.Sa = None

    def parainit(self, y, nu, lambdam_init, lambda0_init):
        """:11-22.  lambdam ~ U(.2,1), lambda0 ~ U(.1,.7) are inputs."""
        self.parameters = OrderedDict(
            sigma=math.log(r_var(y)),
            lambdam=math.log(lambdam_init),
            nu=math.log(nu),
            lambda0=math.log(lambda0_init),
            Lm=np.full(self.p, -0.1),
            La=np.full(self.p, 0.1),
            mu=float(np.mean(y)),
        )

    def kernel_mat(self, X1, X2, z1, z2):  # :23-31
        return kernmat_TP_SE_cpp(_as_matrix(X1), _as_matrix(X2), z1, z2, self.parameters)

    def getinv_kernel(self, X, z):  # :32-45
        Sl = self.kernel_mat(X, X, z, z)
        self.Smat, self.Sm, self.Sa = Sl["Smat"], Sl["Sm"], Sl["Sa"]
        self.invSmatn = invkernel_cpp(z, self.w, self.Smat, self.parameters)

    def para_update(self, it, y, X, z, Optim, faithful=True):  # :46-60 (no get_train_stats)
        self.getinv_kernel(X, z)
        gl = grad_TP_SE_cpp(y, _as_matrix(X), z, self.w, self.Smat, self.Sm, self.Sa, self.invSmatn, self.parameters)
        self.parameters = Optim.update(it, self.parameters, gl["gradients"])
        self.mean_solution(y, z)
        return gl["stats"]

    def get_train_stats(self, y, X, z):  # :61-66
        self.getinv_kernel(X, z)
        return stats_TP_SE(y, self.Smat, self.invSmatn, self.parameters)

    def mean_solution(self, y, z):  # :67-72
        self.parameters["mu"] = mu_solution_cpp(y, z, self.invSmatn, self.parameters)

    def _sample_U(self, cov, rng, want_ate=False):
        n2 = cov.shape[0]
        U = np.zeros(n2)
        ate_U = 0.0
        N_sample = 1000
        N_rep = int(math.ceil(self.nsampling / N_sample))  # :88
        df = math.exp(self.parameters["nu"])  # `$n` partial-matches nu (:89)
        for _ in range(N_rep):
            S = rmvt(N_sample, cov, df, rng)
            U += (1.0 / N_rep) * np.sort(S, axis=0)[int(math.floor(N_sample * 0.90)) - 1, :]  # :90 (1-based)
            if want_ate:
                ate_U += (1.0 / N_rep) * np.sort(S.mean(axis=1))[int(math.floor(N_sample * 0.90)) - 1]  # :121
        return U, ate_U

    def predict(self, y, X, z, X2, z2, rng=None):  # :72-94
        X2 = _as_matrix(X2)
        n2 = X2.shape[0]
        S_xX = self.kernel_mat(X2, X, z2, z)["Smat"]
        S_xx = self.kernel_mat(X2, X2, z2, z2)["Smat"]
        mu = self.parameters["mu"]
        map_ = S_xX @ self.invSmatn @ (np.asarray(y) - mu) + mu  # :78
        cov = S_xx - S_xX @ self.invSmatn @ S_xX.T + math.exp(self.parameters["sigma"]) * np.eye(n2)  # :79
        out = dict(map=map_, cov=cov)
        if rng is not None:
            U, _ = self._sample_U(cov, rng)
            out["ci"] = np.column_stack([-U, U])  # :93
        return out

    def predict_treat(self, y, X, z, X2, rng=None):  # :95-125
        X2 = _as_matrix(X2)
        n2 = X2.shape[0]
        one = np.ones(n2)
        Sa_xX = self.kernel_mat(X2, X, one, z)["Sa"]
        Sa_xx = self.kernel_mat(X2, X2, one, one)["Sa"]
        map_ = Sa_xX @ self.invSmatn @ (np.asarray(y) - self.parameters["mu"])  # :107
        ate = float(np.mean(map_))
        cov = Sa_xx - Sa_xX @ self.invSmatn @ Sa_xX.T + 1e-6 * np.eye(n2)  # :110
        out = dict(map=map_, ate_map=ate, cov=cov)
        if rng is not None:
            U, ate_U = self._sample_U(cov, rng, want_ate=True)
            out["ci"] = np.column_stack([-U, U])
            out["ate_ci"] = np.array([-ate_U, ate_U])  # :124
        return out


# --------------------------------------------------------------------------
# R/main.R, R/predict.CausalStump.R, R/treatment.R
# --------------------------------------------------------------------------
class CausalStumpFit:
    def __init__(self, Kernel, moments, train_data, stats, iters):
        self.Kernel, self.moments, self.train_data = Kernel, moments, train_data
        self.stats, self.iters = stats, iters


def CausalStump(y, X, z, w=None, pscore=None, myoptim="Nadam", maxiter=5000, tol=1e-4, prior=False, nu=200,
                nsampling=5000, learning_rate=0.01, beta1=0.9, beta2=0.999, momentum=0.0,
                init_draws=(0.6, 0.6), faithful=True):
    """R/main.R:35-96.  ``init_draws`` replaces the two runif draws of parainit
    (GP: lambdam, lambdaa ~ U(.2,1); TP: lambdam ~ U(.2,1), lambda0 ~ U(.1,.7)).
    ``faithful=False`` skips the redundant per-iteration get_train_stats of the GP
    class (R/kernel_GP_SE_class.R:58), which does not change any returned number.
    """
    X = _as_matrix(X)
    if pscore is not None:
        X = np.column_stack([X, np.asarray(pscore, dtype=np.float64)])  # :38
    n, p = len(y), X.shape[1]
    nr = norm_variables(y, X)  # :43
    moments, y, X = nr["moments"], nr["y"], nr["X"]
    z = np.asarray(z, dtype=np.float64)
    if w is None:
        w = np.ones(n)  # :47
    if prior:
        K = KernelClass_TP_SE(p, w, nsampling)
        K.parainit(y, nu, *init_draws)
    else:
        K = KernelClass_GP_SE(p, w)
        K.parainit(y, *init_draws)
    if myoptim == "Adam":
        opt = optAdam(learning_rate, beta1, beta2)
    elif myoptim == "Nadam":
        opt = optNadam(learning_rate, beta1, beta2)
    elif myoptim in ("GD", "Nesterov"):
        if myoptim == "GD":
            momentum = 0.0  # :69
        opt = optNestorov(learning_rate, momentum)
    else:
        raise ValueError(myoptim)
    opt.initOpt(K)  # :74
    stats = np.zeros((2, maxiter + 2))  # :77
    it = 0
    for it in range(1, maxiter + 1):  # :79-84
        stats[:, it] = K.para_update(it, y, X, z, opt, faithful=faithful)
        change = abs(stats[1, it] - stats[1, it - 1])
        if change < tol and it > 3:
            break
    stats[:, it + 1] = K.get_train_stats(y, X, z)  # :88
    return CausalStumpFit(K, moments, dict(y=y, X=X, z=z), stats[:, : it + 2], it)


def predict(cs, X=None, z=None, pscore=None, nsampling=None, rng=None):
    """R/predict.CausalStump.R:10-39."""
    if X is None:
        Xn = cs.train_data["X"]
    else:
        X = _as_matrix(X)
        if pscore is not None:
            X = np.column_stack([X, pscore])
        if X.shape[1] != cs.train_data["X"].shape[1]:
            raise ValueError("Error: propensity score mismatch")
        Xn = norm_variables(X=X, moments=cs.moments)["X"]  # :18
    n = Xn.shape[0]
    if z is None:
        z = cs.train_data["z"]
    elif np.ndim(z) == 0:
        z = np.full(n, float(z))
    if nsampling is not None and cs.Kernel.kind == "TP":
        cs.Kernel.nsampling = nsampling
    td = cs.train_data
    if cs.Kernel.kind == "TP":
        pl = cs.Kernel.predict(td["y"], td["X"], td["z"], Xn, np.asarray(z, dtype=np.float64), rng=rng)
    else:
        pl = cs.Kernel.predict(td["y"], td["X"], td["z"], Xn, np.asarray(z, dtype=np.float64))
    sdY = math.sqrt(cs.moments["varY"])
    map_ = cs.moments["meanY"] + sdY * pl["map"]  # :35
    out = dict(map=map_)
    if "ci" in pl:
        out["ci"] = cs.moments["varY"] * pl["ci"] + map_[:, None]  # :36
    return out


def treatment(cs, X=None, pscore=None, nsampling=None, rng=None):
    """R/treatment.R:3-34."""
    if X is None:
        Xn = cs.train_data["X"]
    else:
        X = _as_matrix(X)
        if pscore is not None:
            X = np.column_stack([X, pscore])
        if X.shape[1] != cs.train_data["X"].shape[1]:
            raise ValueError("Error: propensity score mismatch")
        Xn = norm_variables(X=X, moments=cs.moments)["X"]
    if nsampling is not None and cs.Kernel.kind == "TP":
        cs.Kernel.nsampling = nsampling
    td = cs.train_data
    if cs.Kernel.kind == "TP":
        pl = cs.Kernel.predict_treat(td["y"], td["X"], td["z"], Xn, rng=rng)
    else:
        pl = cs.Kernel.predict_treat(td["y"], td["X"], td["z"], Xn)
    sdY = math.sqrt(cs.moments["varY"])
    varY = cs.moments["varY"]
    map_ = sdY * pl["map"]  # :28
    out = dict(map=map_, ate=sdY * pl["ate_map"])
    if "ci" in pl:
        out["ci"] = varY * pl["ci"] + map_[:, None]  # :29
    if "ate_ci" in pl:
        out["ate_ci"] = varY * pl["ate_ci"] + out["ate"]  # :31
    return out


# --------------------------------------------------------------------------
# flat parameter-vector helpers (layout of the C-ABI, include/cs_b200.h)
# --------------------------------------------------------------------------
def pack_params(par):
    """dict -> flat vector in reference list order (scalars then Lm[p], La[p], mu)."""
    out = []
    for v in par.values():
        out.extend(np.atleast_1d(np.asarray(v, dtype=np.float64)).tolist())
    return np.array(out, dtype=np.float64)


def unpack_params(vec, kind, p):
    order = GP_ORDER if kind == "GP" else TP_ORDER
    out, k = OrderedDict(), 0
    for name in order:
        if name in ("Lm", "La"):
            out[name] = np.array(vec[k:k + p], dtype=np.float64)
            k += p
        else:
            out[name] = float(vec[k])
            k += 1
    return out


def eval_lml_grad(kind, y, X, z, w, par):
    """One 'LML + gradient evaluation' (SURVEY.md 8d): Gram + inverse + grad/stats
    + mu solution, reference-faithful arithmetic.  Returns (grad dict, stats, mu_sol)."""
    X = _as_matrix(X)
    if kind == "GP":
        Kl = kernmat_GP_SE_cpp(X, X, z, z, par)
        inv = invkernel_cpp(z, w, Kl["Kmat"], par)
        gl = grad_GP_SE_cpp(y, X, z, w, Kl["Kmat"], Kl["Km"], Kl["Ka"], inv, par)
    else:
        Kl = kernmat_TP_SE_cpp(X, X, z, z, par)
        inv = invkernel_cpp(z, w, Kl["Smat"], par)
        gl = grad_TP_SE_cpp(y, X, z, w, Kl["Smat"], Kl["Sm"], Kl["Sa"], inv, par)
    mu_sol = mu_solution_cpp(y, z, inv, par)
    return gl["gradients"], gl["stats"], mu_sol
