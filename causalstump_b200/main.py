"""Host-side mirror of the reference's R API on top of the device-resident fit handle.

  CausalStump(y, X, z, ...)     /root/reference/R/main.R:35-96
  predict(cs, X, z, ...)        R/predict.CausalStump.R:10-39
  treatment(cs, X, ...)         R/treatment.R:3-34
  KernelClass_GP_SE / _TP_SE    R/kernel_GP_SE_class.R, R/kernel_TP_SE_class.R
  optAdam / optNadam / optNestorov, list2zero   R/optimizer_classes.R
  norm_variables, check_inputs  R/utilities.R

Same names, argument meaning and error behaviour.  All numerics of the hot path run on
the GPU through the C ABI (include/cs_b200.h); the optimisation loop of R/main.R:79-84
is one device-side loop (cs_fit_run).  Host code here is O(n p) data preparation only.
The reference's R toolchain is absent from the build image, which is why this mirror is
Python; the R-side binding is in causalstump_b200/r/ and INTEGRATION.md.
"""
from __future__ import annotations

import math
import warnings
from collections import OrderedDict

import numpy as np

from . import _lib, rcpp_exports as rx


# --------------------------------------------------------------------------------------
# R/utilities.R
# --------------------------------------------------------------------------------------
def _r_var(x):
    return float(np.var(np.asarray(x, dtype=np.float64), ddof=1))


def norm_variables(y=None, X=None, moments=None):
    """R/utilities.R:1-33 (quirk Q9: max-abs of the uncentred column; y not centred)."""
    X = np.array(X, dtype=np.float64, copy=True)
    if X.ndim == 1:
        X = X[:, None]
    p = X.shape[1]
    if moments is None:
        moments = dict(meanX=np.zeros(p), varX=np.ones(p), meanY=0.0, varY=_r_var(y))
        for j in range(p):
            col = X[:, j]
            col = col[~np.isnan(col)]
            binary = bool(np.all((col == 0.0) | (col == 1.0)))
            unit = bool(np.all((col <= 1.0) & (col >= 0.0)))
            if not binary and not unit:
                moments["meanX"][j] = col.mean()
                moments["varX"][j] = np.max(np.abs(X[:, j])) ** 2
    Xn = (X - moments["meanX"][None, :]) / np.sqrt(moments["varX"])[None, :]
    if y is None:
        return dict(X=Xn)
    yn = (np.asarray(y, dtype=np.float64) - moments["meanY"]) / math.sqrt(moments["varY"])
    return dict(moments=moments, y=yn, X=Xn)


def check_inputs(y, X, z):
    """R/utilities.R:35-45 -- warnings only."""
    X = np.asarray(X)
    if len(y) != X.shape[0]:
        warnings.warn("y and X: nr of observations mismatch")
    if len(y) != len(z):
        warnings.warn("y and z: nr of observations mismatch")


# --------------------------------------------------------------------------------------
# R/optimizer_classes.R
# --------------------------------------------------------------------------------------
def list2zero(mylist):
    return OrderedDict((k, 0.0 if np.ndim(v) == 0 else np.zeros(len(v))) for k, v in mylist.items())


class optAdam:
    code = _lib.CS_OPT_ADAM

    def __init__(self, lr, beta1, beta2):
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2
        self.momentum = 0.0
        self.m = self.v = None

    def initOpt(self, KernelObj):
        self.m = list2zero(KernelObj.parameters)
        self.v = list2zero(KernelObj.parameters)

    def update(self, iter, parameters, gradients):  # noqa: A002
        out = rx.Adam_cpp(iter, self.lr, self.beta1, self.beta2, 1e-8, self.m, self.v, gradients, parameters)
        self.m, self.v = out["m"], out["v"]
        return out["parameters"]


class optNadam(optAdam):
    code = _lib.CS_OPT_NADAM

    def update(self, iter, parameters, gradients):  # noqa: A002
        out = rx.Nadam_cpp(iter, self.lr, self.beta1, self.beta2, 1e-8, self.m, self.v, gradients, parameters)
        self.m, self.v = out["m"], out["v"]
        return out["parameters"]


class optNestorov:
    code = _lib.CS_OPT_NESTEROV

    def __init__(self, lr, momentum):
        self.lr, self.momentum = lr, momentum
        self.beta1 = self.beta2 = 0.0
        self.nu = None

    def initOpt(self, KernelObj):
        self.nu = list2zero(KernelObj.parameters)

    def update(self, iter, parameters, gradients):  # noqa: A002
        out = rx.Nesterov_cpp(self.lr, self.momentum, self.nu, gradients, parameters)
        self.nu = out["nu"]
        return out["parameters"]


# --------------------------------------------------------------------------------------
# kernel classes
# --------------------------------------------------------------------------------------
class _KernelBase:
    kind = _lib.CS_GP
    _names = ("Kmat", "Km", "Ka", "invKmatn")

    def __init__(self, p, w):
        self.p = int(p)
        self.w = np.asarray(w, dtype=np.float64)
        self.parameters = None
        self._fit = None       # device handle (cs_fit)
        self._fit_key = None

    # -- device handle ------------------------------------------------------------
    @staticmethod
    def _fingerprint(*arrays):
        """Cheap content check (Adler-32 of the bytes, ~1 GB/s): catches in-place edits of y, X or z between calls."""
        import zlib
        h = 1
        for a in arrays:
            h = zlib.adler32(np.ascontiguousarray(a).view(np.uint8), h)
        return h

    def invalidate(self):
        """Drop the device handle: the next call re-uploads y, X, z (call after changing them in place if the
        content check below is bypassed, or to release the device memory)."""
        if self._fit is not None:
            self._fit.close()
        self._fit, self._fit_key = None, None

    def _handle(self, y, X, z, Optim=None, need_inverse=False):
        # The cache entry keeps STRONG references to the caller's arrays (identity compared with `is`, so a freed
        # array whose id() is reused can never alias) plus a content fingerprint (so an in-place edit re-uploads).
        # X is keyed on the object the caller passed: a 1-D X gets a fresh 2-D view per call, which is not what is keyed.
        src = (y, X, z)
        Xm = rx._mat(X)
        fp = (Xm.shape, self._fingerprint(np.asarray(y, dtype=np.float64), Xm, np.asarray(z, dtype=np.float64)))
        k = self._fit_key
        fresh = self._fit is None or k is None or not (k[0][0] is src[0] and k[0][1] is src[1] and k[0][2] is src[2]) or k[1] != fp
        key = (src, fp)
        X = Xm
        if fresh:
            if self._fit is not None:
                self._fit.close()
            kw = {}
            if Optim is not None:
                kw = dict(optim=Optim.code, lr=Optim.lr, beta1=Optim.beta1, beta2=Optim.beta2, momentum=Optim.momentum)
            self._fit = _lib.Fit(rx.lib(), self.kind, y, X, z, self.w, rx.pack(self.parameters), **kw)
            self._fit_key = key
            if need_inverse:  # a new handle has no K^-1 yet: build it for the current parameters
                self._fit.eval()
        return self._fit

    def _matrices(self, which):
        """Lazy download of the RefClass matrix fields (R/kernel_GP_SE_class.R:2-8)."""
        if self._fit is None:
            raise RuntimeError("kernel has no device state yet: call getinv_kernel / para_update first")
        m = self._fit.get_matrices(want=(which,))
        return m[which]

    @property
    def invKmatn(self):
        return self._matrices("invK")

    @property
    def Kmat(self):
        return self._matrices("Kmat")

    @property
    def Km(self):
        return self._matrices("Km")

    @property
    def Ka(self):
        return self._matrices("Ka")

    # -- reference methods --------------------------------------------------------
    def getinv_kernel(self, X, z, y=None):
        """R/kernel_GP_SE_class.R:31-44: Gram + noisy inverse for the current parameters
        (device-resident; the matrix fields are materialised lazily)."""
        if y is None:
            y = np.zeros(rx._mat(X).shape[0])
        f = self._handle(y, X, z)
        f.set_params(rx.pack(self.parameters))
        f.eval()

    def para_update(self, iter, y, X, z, Optim):  # noqa: A002
        """One host-driven iteration (R/kernel_GP_SE_class.R:45-60).  CausalStump() does
        not use this: it runs the whole loop on the device (cs_fit_run)."""
        f = self._handle(y, X, z, Optim)
        g, stats, mu_sol = f.eval(rx.pack(self.parameters))
        gradients = rx.unpack(g, self.parameters)
        self.parameters = Optim.update(iter, self.parameters, gradients)
        self.parameters["mu"] = mu_sol  # mean_solution with the pre-update inverse (quirk Q1)
        if iter % 100 == 0:
            print("%5d | log Evidence %9.4f | RMSE %9.4f " % (iter, stats[1], stats[0]))
        return stats

    def get_train_stats(self, y, X, z):
        f = self._handle(y, X, z)
        _, stats, _ = f.eval(rx.pack(self.parameters))
        return stats

    def mean_solution(self, y, z):
        _, _, mu = self._fit.fetch()
        self.parameters["mu"] = mu


class KernelClass_GP_SE(_KernelBase):
    """R/kernel_GP_SE_class.R:1-107."""
    kind = _lib.CS_GP

    def parainit(self, y, draws):
        self.parameters = OrderedDict(
            sigma=math.log(_r_var(y)), lambdam=math.log(draws[0]), lambdaa=math.log(draws[1]),
            Lm=np.full(self.p, -0.1), La=np.full(self.p, 0.1), mu=float(np.mean(y)))

    def kernel_mat(self, X1, X2, z1, z2):
        return rx.kernmat_GP_SE_cpp(rx._mat(X1), rx._mat(X2), z1, z2, self.parameters)

    def predict(self, y, X, z, X2, z2):
        """:74-85; ci = +-1.96 * variance (quirk Q8)."""
        r = self._handle(y, X, z, need_inverse=True).predict(X2, z2)
        return dict(map=r["map"], ci=np.column_stack([-1.96 * r["var"], 1.96 * r["var"]]))

    def predict_treat(self, y, X, z, X2):
        """:86-105."""
        r = self._handle(y, X, z, need_inverse=True).treat(X2)
        s = 1.96 * math.sqrt(r["ate_var"])
        return dict(map=r["map"], ci=np.column_stack([-1.96 * r["var"], 1.96 * r["var"]]), ate_map=r["ate"],
                    ate_ci=np.array([-s, s]))


class KernelClass_TP_SE(_KernelBase):
    """R/kernel_TP_SE_class.R:1-127."""
    kind = _lib.CS_TP
    _names = ("Smat", "Sm", "Sa", "invSmatn")

    def __init__(self, p, w, nsampling=5000):
        super().__init__(p, w)
        self.nsampling = nsampling
        self.seed = 20170905

    invSmatn = _KernelBase.invKmatn
    Smat = _KernelBase.Kmat
    Sm = _KernelBase.Km
    Sa = _KernelBase.Ka

    def parainit(self, y, nu, draws):
        self.parameters = OrderedDict(
            sigma=math.log(_r_var(y)), lambdam=math.log(draws[0]), nu=math.log(nu), lambda0=math.log(draws[1]),
            Lm=np.full(self.p, -0.1), La=np.full(self.p, 0.1), mu=float(np.mean(y)))

    def kernel_mat(self, X1, X2, z1, z2):
        return rx.kernmat_TP_SE_cpp(rx._mat(X1), rx._mat(X2), z1, z2, self.parameters)

    def predict(self, y, X, z, X2, z2):
        """:72-94: map, then rmvt sampling for the one-sided 90% bound U."""
        r = self._handle(y, X, z, need_inverse=True).predict(X2, z2, want_cov=True)
        U = rx.lib().tp_sample(r["cov"], math.exp(self.parameters["nu"]), self.nsampling, self.seed)
        return dict(map=r["map"], ci=np.column_stack([-U, U]))

    def predict_treat(self, y, X, z, X2):
        """:95-125 (cov + 1e-6 I, :110)."""
        r = self._handle(y, X, z, need_inverse=True).treat(X2, cov_jitter=1e-6, want_cov=True)
        U, aU = rx.lib().tp_sample(r["cov"], math.exp(self.parameters["nu"]), self.nsampling, self.seed + 1,
                                   want_ate=True)
        return dict(map=r["map"], ci=np.column_stack([-U, U]), ate_map=r["ate"], ate_ci=np.array([-aU, aU]))


# --------------------------------------------------------------------------------------
# R/main.R, R/predict.CausalStump.R, R/treatment.R
# --------------------------------------------------------------------------------------
class CausalStumpObject:
    """structure(list(Kernel, moments, train_data), class='CausalStump'), R/main.R:95."""

    def __init__(self, Kernel, moments, train_data, stats, iters):
        self.Kernel, self.moments, self.train_data = Kernel, moments, train_data
        self.stats, self.iters = stats, iters


def CausalStump(y, X, z, w=None, pscore=None, kernelfun="SE", myoptim="Nadam", maxiter=5000, tol=1e-4, prior=False,
                nu=200, nsampling=5000, learning_rate=0.01, beta1=0.9, beta2=0.999, momentum=0.0, init_draws=None,
                rng=None, verbose=True):
    """R/main.R:35-96.  `init_draws` (or `rng`) supplies the two runif draws of parainit
    that R takes from its global RNG (GP: lambdam, lambdaa ~ U(.2,1); TP: lambdam ~
    U(.2,1), lambda0 ~ U(.1,.7))."""
    check_inputs(y, X, z)
    X = rx._mat(X)
    if pscore is not None:
        X = np.column_stack([X, np.asarray(pscore, dtype=np.float64)])
    n, p = len(y), X.shape[1]
    nr = norm_variables(y, X)
    moments, y, X = nr["moments"], np.ascontiguousarray(nr["y"]), np.asfortranarray(nr["X"])
    z = np.ascontiguousarray(np.asarray(z, dtype=np.float64))
    if w is None:
        w = np.ones(n)
    if init_draws is None:
        rng = rng if rng is not None else np.random.default_rng()
        init_draws = (rng.uniform(0.2, 1.0), rng.uniform(0.1, 0.7) if prior else rng.uniform(0.2, 1.0))
    if prior:
        if verbose:
            print("\nFitting the Student-t Process with nu=%d:" % nu)
        myKernel = KernelClass_TP_SE(p=p, w=w, nsampling=nsampling)
        myKernel.parainit(y, nu, init_draws)
    else:
        if verbose:
            print("\nFitting the Gaussian process:")
        myKernel = KernelClass_GP_SE(p=p, w=w)
        myKernel.parainit(y, init_draws)
    if myoptim == "Adam":
        myOptimizer = optAdam(learning_rate, beta1, beta2)
    elif myoptim == "Nadam":
        myOptimizer = optNadam(learning_rate, beta1, beta2)
    elif myoptim in ("GD", "Nesterov"):
        if myoptim == "GD":
            momentum = 0.0
        myOptimizer = optNestorov(learning_rate, momentum)
    else:
        raise ValueError("unknown optimizer %r" % (myoptim,))
    myOptimizer.initOpt(myKernel)

    # the loop of R/main.R:79-84 plus the final get_train_stats (:88), on the device
    fit = myKernel._handle(y, X, z, myOptimizer)
    fit.set_params(rx.pack(myKernel.parameters))
    fit.reset_optimizer()
    iters, trace = fit.run(maxiter, tol)
    myKernel.parameters = rx.unpack(fit.get_params(), myKernel.parameters)
    if verbose:
        if iters < maxiter:
            print("Stopped: change smaller than tolerance after %d iterations" % iters)
        if iters == maxiter:
            print("Stopped: maximum iterations reached")
    return CausalStumpObject(myKernel, moments, dict(y=y, X=X, z=z), trace, iters)


def _new_X(cs, X, pscore):
    if X is None:
        return cs.train_data["X"]
    X = rx._mat(X)
    if pscore is not None:
        X = np.column_stack([X, np.asarray(pscore, dtype=np.float64)])
    if X.shape[1] != cs.train_data["X"].shape[1]:
        raise ValueError("Error: propensity score mismatch")
    return norm_variables(X=X, moments=cs.moments)["X"]


def predict(cs_object, X=None, z=None, pscore=None, nsampling=None):
    """R/predict.CausalStump.R:10-39."""
    Xn = _new_X(cs_object, X, pscore)
    n = Xn.shape[0]
    if z is None:
        z = cs_object.train_data["z"]
    elif np.ndim(z) == 0 or len(np.atleast_1d(z)) == 1:
        z = np.full(n, float(np.atleast_1d(z)[0]))
    if nsampling is not None and isinstance(cs_object.Kernel, KernelClass_TP_SE):
        cs_object.Kernel.nsampling = nsampling
    td = cs_object.train_data
    pl = cs_object.Kernel.predict(td["y"], td["X"], td["z"], Xn, np.asarray(z, dtype=np.float64))
    map_ = cs_object.moments["meanY"] + math.sqrt(cs_object.moments["varY"]) * pl["map"]
    ci = cs_object.moments["varY"] * pl["ci"] + map_[:, None]
    return dict(map=map_, ci=ci)


def treatment(cs_object, X=None, pscore=None, nsampling=None):
    """R/treatment.R:3-34."""
    Xn = _new_X(cs_object, X, pscore)
    if nsampling is not None and isinstance(cs_object.Kernel, KernelClass_TP_SE):
        cs_object.Kernel.nsampling = nsampling
    td = cs_object.train_data
    pl = cs_object.Kernel.predict_treat(td["y"], td["X"], td["z"], Xn)
    sd, var = math.sqrt(cs_object.moments["varY"]), cs_object.moments["varY"]
    map_ = sd * pl["map"]
    ci = var * pl["ci"] + map_[:, None]
    cate = sd * pl["ate_map"]
    cate_ci = var * pl["ate_ci"] + cate
    return dict(map=map_, ci=ci, ate=cate, ate_ci=cate_ci)
