"""ctypes binding of the C ABI declared in include/cs_b200.h.

The product library is ``csrc/libcausalstump_b200.so`` (nvcc, sm_100a).  It is loaded
lazily and there is NO fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libcausalstump_b200.so")

CS_GP, CS_TP = 0, 1
CS_OPT_ADAM, CS_OPT_NADAM, CS_OPT_NESTEROV = 0, 1, 2
CS_NPHASE = 7
CS_ERR_NONFINITE = -6
CS_MAX_P = 128
PHASES = ("gram", "potrf", "trtri", "xtx", "matvec", "grad", "final")

dp = C.POINTER(C.c_double)


class cs_fit_config(C.Structure):
    _fields_ = [("kind", C.c_int), ("optim", C.c_int), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("momentum", C.c_double), ("tile", C.c_int),
                ("chunk", C.c_int), ("use_graph", C.c_int), ("gemm_tile", C.c_int), ("inv_pipeline", C.c_int),
                ("partition", C.c_int)]


class cs_sim_config(C.Structure):
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("test_frac", C.c_double), ("maxiter", C.c_int), ("tol", C.c_double),
                ("lr", C.c_double), ("optim", C.c_int), ("batch", C.c_int), ("seed", C.c_ulonglong)]


CS_SIM_NMETRIC = 22


class CsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cs_b200 error {code}: {msg}")
        self.code = code


class NotPositiveDefinite(CsError):
    """status > 0: arma::inv_sympd would have thrown (src/kernel_GP_SE_cpp.cpp:56)."""


# every exported symbol of include/cs_b200.h: name -> (restype, argtypes)
_SIGS = {
    "cs_version": (C.c_int, []),
    "cs_last_error": (C.c_char_p, []),
    "cs_device_count": (C.c_int, []),
    "cs_set_device": (C.c_int, [C.c_int]),
    "cs_kernmat": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp]),
    "cs_invkernel": (C.c_int, [C.c_int, dp, dp, C.c_double, dp, dp]),
    "cs_mu_solution": (C.c_int, [C.c_int, dp, dp, dp]),
    "cs_grad": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
    "cs_stats": (C.c_int, [C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp]),
    "cs_opt_step": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp,
                              dp, dp]),
    "cs_fit_create": (C.c_int, [C.POINTER(cs_fit_config), C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                C.POINTER(C.c_void_p)]),
    "cs_fit_create_batch": (C.c_int, [C.POINTER(cs_fit_config), C.c_int, C.c_int, C.c_int, C.POINTER(dp), C.POINTER(dp),
                                      C.POINTER(dp), C.POINTER(dp), C.POINTER(dp), C.POINTER(C.c_void_p)]),
    "cs_fit_batch_size": (C.c_int, [C.c_void_p]),
    "cs_fit_select": (C.c_int, [C.c_void_p, C.c_int]),
    "cs_fit_destroy": (None, [C.c_void_p]),
    "cs_fit_nparam": (C.c_int, [C.c_void_p]),
    "cs_fit_eval": (C.c_int, [C.c_void_p, dp, dp, dp, dp]),
    "cs_fit_eval_async": (C.c_int, [C.c_void_p]),
    "cs_fit_sync": (C.c_int, [C.c_void_p]),
    "cs_fit_fetch": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "cs_fit_run": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), dp]),
    "cs_fit_run_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int),
                                  C.POINTER(dp)]),
    "cs_fit_member_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "cs_release_scratch": (C.c_int, []),
    "cs_sim_run": (C.c_int, [C.POINTER(cs_sim_config), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp, dp, C.POINTER(C.c_int), dp,
                             C.POINTER(C.c_int)]),
    "cs_sim_generate": (C.c_int, [C.POINTER(cs_sim_config), C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int), dp, dp, dp,
                                  dp, dp, dp]),
    "cs_fit_set_data": (C.c_int, [C.c_void_p, dp, dp, dp, dp]),
    "cs_fit_get_params": (C.c_int, [C.c_void_p, dp]),
    "cs_fit_set_params": (C.c_int, [C.c_void_p, dp]),
    "cs_fit_reset_optimizer": (C.c_int, [C.c_void_p]),
    "cs_fit_get_matrices": (C.c_int, [C.c_void_p, dp, dp, dp, dp]),
    "cs_predict": (C.c_int, [C.c_void_p, C.c_int, dp, dp, dp, dp, dp]),
    "cs_treat": (C.c_int, [C.c_void_p, C.c_int, dp, C.c_double, dp, dp, dp, dp, dp]),
    "cs_tp_sample": (C.c_int, [C.c_int, dp, C.c_double, C.c_int, C.c_ulonglong, dp, dp]),
    "cs_dist_unique_id": (C.c_int, [C.c_char_p]),
    "cs_dist_create": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_char_p, C.POINTER(C.c_void_p)]),
    "cs_dist_eval": (C.c_int, [C.c_void_p, dp, dp, dp, dp, C.POINTER(C.c_float)]),
    "cs_dist_destroy": (None, [C.c_void_p]),
    "cs_dist_sim_eval": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp]),
    "cs_fit_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "cs_fit_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "cs_fit_eval_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "cs_fit_dense_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "cs_fit_dense_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "cs_debug_check_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cs_debug_timeline": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "cs_launch_count": (C.c_longlong, []),
    "cs_flush_l2": (C.c_int, [C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(dp)


class CsLib:
    """One loaded copy of the C-ABI library."""

    def __init__(self, path=DEFAULT_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -m causalstump_b200.csrc.build` "
                "(nvcc, sm_100a).  causalstump_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.dll, name)  # raises AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args

    # -- helpers -----------------------------------------------------------------
    def check(self, code):
        if code == 0:
            return
        msg = (self.dll.cs_last_error() or b"").decode("utf-8", "replace")
        if code > 0:
            raise NotPositiveDefinite(code, msg)
        raise CsError(code, msg)

    def require_device(self):
        if self.dll.cs_device_count() <= 0:
            raise CsError(-4, "no CUDA device available: causalstump_b200 has no CPU fallback")

    # -- stateless mirrors -----------------------------------------------------------
    def kernmat(self, X1, X2, Z1, Z2, Lm, La, lambdam, lambda_a, want=("K", "Km", "Ka")):
        X1, X2 = _f64(np.atleast_2d(X1) if np.ndim(X1) > 1 else np.asarray(X1)[:, None]), \
            _f64(np.atleast_2d(X2) if np.ndim(X2) > 1 else np.asarray(X2)[:, None])
        n1, p = X1.shape
        n2 = X2.shape[0]
        Z1, Z2, Lm, La = _f64(Z1), _f64(Z2), _f64(np.atleast_1d(Lm)), _f64(np.atleast_1d(La))
        outs = {k: (np.empty((n1, n2), order="F") if k in want else None) for k in ("K", "Km", "Ka")}
        self.check(self.dll.cs_kernmat(n1, n2, p, _ptr(X1), _ptr(X2), _ptr(Z1), _ptr(Z2), _ptr(Lm), _ptr(La),
                                       float(lambdam), float(lambda_a), _ptr(outs["K"]), _ptr(outs["Km"]),
                                       _ptr(outs["Ka"])))
        return outs

    def invkernel(self, w, K, sigma, want_logdet=False):
        K = _f64(K)
        n = K.shape[0]
        w = None if w is None else _f64(w)
        inv = np.empty((n, n), order="F")
        ld = C.c_double(0.0)
        self.check(self.dll.cs_invkernel(n, _ptr(w), _ptr(K), float(sigma), _ptr(inv),
                                         C.cast(C.byref(ld), dp) if want_logdet else None))
        return (inv, ld.value) if want_logdet else inv

    def mu_solution(self, y, invK):
        y, invK = _f64(y), _f64(invK)
        out = C.c_double(0.0)
        self.check(self.dll.cs_mu_solution(len(y), _ptr(y), _ptr(invK), C.cast(C.byref(out), dp)))
        return out.value

    def grad(self, kind, y, X, z, w, params, invK=None, Kmat=None, Km=None, Ka=None):
        y, z, params = _f64(y), _f64(z), _f64(params)
        X = _f64(X if np.ndim(X) > 1 else np.asarray(X)[:, None])
        n, p = X.shape
        w = None if w is None else _f64(w)
        mats = [None if m is None else _f64(m) for m in (Kmat, Km, Ka, invK)]
        g = np.empty(len(params))
        st = np.empty(2)
        self.check(self.dll.cs_grad(kind, n, p, _ptr(y), _ptr(X), _ptr(z), _ptr(w), _ptr(mats[0]), _ptr(mats[1]),
                                    _ptr(mats[2]), _ptr(mats[3]), _ptr(params), _ptr(g), _ptr(st)))
        return g, st

    def stats(self, kind, y, Kmat, invK, mu, nu=1.0):
        y, Kmat, invK = _f64(y), _f64(Kmat), _f64(invK)
        st = np.empty(2)
        self.check(self.dll.cs_stats(kind, len(y), _ptr(y), _ptr(Kmat), _ptr(invK), float(mu), float(nu), _ptr(st)))
        return st

    def opt_step(self, optim, it, lr, beta1, beta2, eps, m, v, grad, para):
        m, grad, para = _f64(m).copy(), _f64(grad), _f64(para).copy()
        v = None if v is None else _f64(v).copy()
        self.check(self.dll.cs_opt_step(optim, len(para), float(it), float(lr), float(beta1), float(beta2), float(eps),
                                        _ptr(m), _ptr(v), _ptr(grad), _ptr(para)))
        return m, v, para

    def tp_sample(self, cov, df, nsampling, seed, want_ate=False):
        cov = _f64(cov)
        n2 = cov.shape[0]
        U = np.empty(n2)
        a = C.c_double(0.0)
        self.check(self.dll.cs_tp_sample(n2, _ptr(cov), float(df), int(nsampling), int(seed), _ptr(U),
                                         C.cast(C.byref(a), dp) if want_ate else None))
        return (U, a.value) if want_ate else U

    def launch_count(self):
        return int(self.dll.cs_launch_count())

    # -- simulation driver on the device (cs_sim_*) ------------------------------------------
    @staticmethod
    def _sim_cfg(n, p, test_frac, maxiter, tol, lr, optim, batch, seed):
        return cs_sim_config(int(n), int(p), float(test_frac), int(maxiter), float(tol), float(lr), int(optim), int(batch), int(seed))

    def sim_run(self, units, n=747, p=25, X=None, z=None, test_frac=0.1, maxiter=3000, tol=1e-1, lr=0.01, optim=CS_OPT_NADAM,
                batch=32, seed=565):
        """units = [(replicate, restart), ...] -> dict(metrics[nunits, 22], iters, lml, status): generator, split, normalisation,
        fits, predict / treatment and the 22-metric block all on the device (cs_sim_run)."""
        nu = len(units)
        rep = (C.c_int * nu)(*[int(u[0]) for u in units])
        rs = (C.c_int * nu)(*[int(u[1]) for u in units])
        if X is not None:
            X = _f64(X); z = _f64(z)
            n, p = X.shape
        cfg = self._sim_cfg(n, p, test_frac, maxiter, tol, lr, optim, batch, seed)
        met = np.empty((nu, CS_SIM_NMETRIC))
        its, st = (C.c_int * nu)(), (C.c_int * nu)()
        lml = np.empty(nu)
        self.check(self.dll.cs_sim_run(C.byref(cfg), nu, rep, rs, _ptr(X), _ptr(z), _ptr(met), its, _ptr(lml), st))
        return dict(metrics=met, iters=np.array(list(its)), lml=lml, status=np.array(list(st)))

    def sim_generate(self, replicate, restart, n=747, p=25, X=None, z=None, test_frac=0.1, seed=565, kind=CS_GP):
        """The device generator alone for one unit (test hook, cs_sim_generate)."""
        if X is not None:
            X = _f64(X); z = _f64(z)
            n, p = X.shape
        import math
        ntr = n - int(math.ceil(n * test_frac))
        cfg = self._sim_cfg(n, p, test_frac, 1, 0.0, 0.01, CS_OPT_NADAM, 1, seed)
        o = dict(Xraw=np.empty((n, p), order="F"), z=np.empty(n), y=np.empty(n), y1=np.empty(n), y0=np.empty(n),
                 is_test=np.empty(n, dtype=np.int32), Xtrain=np.empty((ntr, p), order="F"), ytrain=np.empty(ntr), ztrain=np.empty(ntr),
                 Xall=np.empty((n, p), order="F"), moments=np.empty(2 * p + 2), params=np.empty(2 * p + 4))
        self.check(self.dll.cs_sim_generate(C.byref(cfg), int(replicate), int(restart), _ptr(X), _ptr(z), _ptr(o["Xraw"]), _ptr(o["z"]),
                                            _ptr(o["y"]), _ptr(o["y1"]), _ptr(o["y0"]), o["is_test"].ctypes.data_as(C.POINTER(C.c_int)),
                                            _ptr(o["Xtrain"]), _ptr(o["ytrain"]), _ptr(o["ztrain"]), _ptr(o["Xall"]), _ptr(o["moments"]),
                                            _ptr(o["params"])))
        return o

    # -- config 5: distributed evaluation -------------------------------------------------
    def dist_sim_eval(self, kind, y, X, z, w, params, P, Q, T=128):
        """All P*Q ranks in this process (collectives = device copies)."""
        y, z, params = _f64(y), _f64(z), _f64(params)
        X = _f64(X if np.ndim(X) > 1 else np.asarray(X)[:, None])
        n, p = X.shape
        w = None if w is None else _f64(w)
        g, st, mu = np.empty(len(params)), np.empty(2), C.c_double(0.0)
        self.check(self.dll.cs_dist_sim_eval(kind, n, p, T, P, Q, _ptr(y), _ptr(X), _ptr(z), _ptr(w), _ptr(params),
                                             _ptr(g), _ptr(st), C.cast(C.byref(mu), dp)))
        return g, st, mu.value

    def dist_unique_id(self):
        buf = C.create_string_buffer(128)
        self.check(self.dll.cs_dist_unique_id(buf))
        return buf.raw


class DistFit:
    """One rank of the P x Q distributed evaluation (cs_dist, NCCL)."""

    def __init__(self, lib: "CsLib", kind, y, X, z, w, params, rank, world, P, Q, unique_id):
        self.lib = lib
        y, z, params = _f64(y), _f64(z), _f64(params)
        X = _f64(X if np.ndim(X) > 1 else np.asarray(X)[:, None])
        n, p = X.shape
        w = None if w is None else _f64(w)
        self.nparam = len(params)
        h = C.c_void_p(None)
        lib.check(lib.dll.cs_dist_create(kind, n, p, _ptr(y), _ptr(X), _ptr(z), _ptr(w), _ptr(params), rank, world, P, Q,
                                         unique_id, C.byref(h)))
        self.h = h

    def eval(self, params=None):
        g, st, mu, ms = np.empty(self.nparam), np.empty(2), C.c_double(0.0), C.c_float(0.0)
        pp = None if params is None else _f64(params)
        self.lib.check(self.lib.dll.cs_dist_eval(self.h, _ptr(pp), _ptr(g), _ptr(st), C.cast(C.byref(mu), dp), C.byref(ms)))
        return g, st, mu.value, ms.value

    def close(self):
        if self.h:
            self.lib.dll.cs_dist_destroy(self.h)
            self.h = C.c_void_p(None)


class Fit:
    """Device-resident fit handle (cs_fit)."""

    def __init__(self, lib: CsLib, kind, y, X, z, w, init_params, optim=CS_OPT_NADAM, lr=0.01, beta1=0.9,
                 beta2=0.999, eps=1e-8, momentum=0.0, tile=0, chunk=0, use_graph=0, gemm_tile=0, inv_pipeline=0, partition=0):
        self.lib = lib
        self.h = C.c_void_p(None)
        y, z, init_params = _f64(y), _f64(z), _f64(init_params)
        X = _f64(X if np.ndim(X) > 1 else np.asarray(X)[:, None])
        self.n, self.p = X.shape
        self.kind = kind
        w = None if w is None else _f64(w)
        cfg = cs_fit_config(kind, optim, lr, beta1, beta2, eps, momentum, tile, chunk, use_graph, gemm_tile, inv_pipeline, partition)
        h = C.c_void_p(None)
        lib.check(lib.dll.cs_fit_create(C.byref(cfg), self.n, self.p, _ptr(y), _ptr(X), _ptr(z), _ptr(w),
                                        _ptr(init_params), C.byref(h)))
        self.h = h
        self.nparam = lib.dll.cs_fit_nparam(self.h)

    def close(self):
        if self.h:
            self.lib.dll.cs_fit_destroy(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self, params=None):
        g, st, mu = np.empty(self.nparam), np.empty(2), C.c_double(0.0)
        pp = None if params is None else _f64(params)
        self.lib.check(self.lib.dll.cs_fit_eval(self.h, _ptr(pp), _ptr(g), _ptr(st), C.cast(C.byref(mu), dp)))
        return g, st, mu.value

    def eval_async(self):
        self.lib.check(self.lib.dll.cs_fit_eval_async(self.h))

    def sync(self):
        self.lib.check(self.lib.dll.cs_fit_sync(self.h))

    def fetch(self):
        g, st, mu = np.empty(self.nparam), np.empty(2), C.c_double(0.0)
        self.lib.check(self.lib.dll.cs_fit_fetch(self.h, _ptr(g), _ptr(st), C.cast(C.byref(mu), dp)))
        return g, st, mu.value

    def run(self, maxiter, tol):
        it = C.c_int(0)
        trace = np.zeros((2, maxiter + 2), order="F")
        self.lib.check(self.lib.dll.cs_fit_run(self.h, int(maxiter), float(tol), C.byref(it), _ptr(trace)))
        return it.value, trace[:, : it.value + 2]

    def get_params(self):
        out = np.empty(self.nparam)
        self.lib.check(self.lib.dll.cs_fit_get_params(self.h, _ptr(out)))
        return out

    def set_params(self, params):
        self.lib.check(self.lib.dll.cs_fit_set_params(self.h, _ptr(_f64(params))))

    def set_data(self, y, X, z, w=None):
        y, z = _f64(y), _f64(z)
        X = _f64(X if np.ndim(X) > 1 else np.asarray(X)[:, None])
        assert X.shape == (self.n, self.p) and len(y) == self.n
        w = None if w is None else _f64(w)
        self.lib.check(self.lib.dll.cs_fit_set_data(self.h, _ptr(y), _ptr(X), _ptr(z), _ptr(w)))

    def reset_optimizer(self):
        self.lib.check(self.lib.dll.cs_fit_reset_optimizer(self.h))

    def get_matrices(self, want=("Kmat", "Km", "Ka", "invK")):
        n = self.n
        outs = {k: (np.empty((n, n), order="F") if k in want else None) for k in ("Kmat", "Km", "Ka", "invK")}
        self.lib.check(self.lib.dll.cs_fit_get_matrices(self.h, _ptr(outs["Kmat"]), _ptr(outs["Km"]),
                                                        _ptr(outs["Ka"]), _ptr(outs["invK"])))
        return outs

    def predict(self, X2, z2, want_cov=False):
        X2 = _f64(X2 if np.ndim(X2) > 1 else np.asarray(X2)[:, None])
        n2 = X2.shape[0]
        z2 = _f64(np.broadcast_to(np.asarray(z2, dtype=np.float64), (n2,)).copy())
        mp, var = np.empty(n2), np.empty(n2)
        cov = np.empty((n2, n2), order="F") if want_cov else None
        self.lib.check(self.lib.dll.cs_predict(self.h, n2, _ptr(X2), _ptr(z2), _ptr(mp), _ptr(var), _ptr(cov)))
        return dict(map=mp, var=var, cov=cov)

    def treat(self, X2, cov_jitter=0.0, want_cov=False):
        X2 = _f64(X2 if np.ndim(X2) > 1 else np.asarray(X2)[:, None])
        n2 = X2.shape[0]
        mp, var = np.empty(n2), np.empty(n2)
        ate, av = C.c_double(0.0), C.c_double(0.0)
        cov = np.empty((n2, n2), order="F") if want_cov else None
        self.lib.check(self.lib.dll.cs_treat(self.h, n2, _ptr(X2), float(cov_jitter), _ptr(mp), _ptr(var),
                                             C.cast(C.byref(ate), dp), C.cast(C.byref(av), dp), _ptr(cov)))
        return dict(map=mp, var=var, ate=ate.value, ate_var=av.value, cov=cov)

    def event_record(self, slot):
        self.lib.check(self.lib.dll.cs_fit_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float(0.0)
        self.lib.check(self.lib.dll.cs_fit_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def eval_profile(self):
        arr = (C.c_float * CS_NPHASE)()
        self.lib.check(self.lib.dll.cs_fit_eval_profile(self.h, arr))
        return dict(zip(PHASES, [float(x) for x in arr]))

    def dense_timing(self, on=True):
        self.lib.check(self.lib.dll.cs_fit_dense_timing(self.h, 1 if on else 0))

    def dense_ms(self):
        """(summed ms, evaluations) of the dense phase since dense_timing(True) / the last call"""
        ms, k = C.c_float(0.0), C.c_int(0)
        self.lib.check(self.lib.dll.cs_fit_dense_ms(self.h, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def timeline(self, cap=4096):
        """Per-launch timeline of the dense plan: rows (stream, type, tiles, K|block, t0_ms, t1_ms)."""
        buf = np.zeros((cap, 6), dtype=np.float32)
        n = C.c_int(0)
        self.lib.check(self.lib.dll.cs_debug_timeline(self.h, cap, C.byref(n), buf.ctypes.data_as(C.POINTER(C.c_float))))
        return buf[: n.value].copy()

    def flush_l2(self):
        self.lib.check(self.lib.dll.cs_flush_l2(self.h))

    def member_status(self):
        st = (C.c_int * max(getattr(self, "B", 1), 1))()
        self.lib.check(self.lib.dll.cs_fit_member_status(self.h, st))
        return [int(x) for x in st]


class FitBatch(Fit):
    """Batched handle (cs_fit_create_batch): B independent fits of the same (n, p, kind) advance with every
    launch.  `members` is a list of (y, X, z, w, init_params).  eval_async / run act on all members;
    predict / treat / get_params / fetch / get_matrices act on the member chosen with select(b)."""

    def __init__(self, lib: CsLib, kind, members, optim=CS_OPT_NADAM, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8,
                 momentum=0.0, tile=0, chunk=0, use_graph=0, gemm_tile=0):
        self.lib = lib
        self.h = C.c_void_p(None)
        B = len(members)
        assert B >= 1
        ys, Xs, zs, ws, ps = [], [], [], [], []
        for (y, X, z, w, ip) in members:
            ys.append(_f64(y)); zs.append(_f64(z)); ps.append(_f64(ip))
            Xs.append(_f64(X if np.ndim(X) > 1 else np.asarray(X)[:, None]))
            ws.append(None if w is None else _f64(w))
        self.n, self.p = Xs[0].shape
        assert all(X.shape == (self.n, self.p) for X in Xs) and all(len(y) == self.n for y in ys)
        self.kind, self.B = kind, B
        arr = lambda lst: (dp * B)(*[_ptr(a) for a in lst])
        cfg = cs_fit_config(kind, optim, lr, beta1, beta2, eps, momentum, tile, chunk, use_graph, gemm_tile, 0, 0)
        h = C.c_void_p(None)
        lib.check(lib.dll.cs_fit_create_batch(C.byref(cfg), B, self.n, self.p, arr(ys), arr(Xs), arr(zs), arr(ws), arr(ps),
                                              C.byref(h)))
        self.h = h
        self.nparam = lib.dll.cs_fit_nparam(self.h)

    def select(self, member):
        self.lib.check(self.lib.dll.cs_fit_select(self.h, int(member)))
        return self

    def member_status(self):
        """Per member after run(): 0 OK, > 0 failing pivot (not positive definite), CS_ERR_NONFINITE (-6)."""
        st = (C.c_int * self.B)()
        self.lib.check(self.lib.dll.cs_fit_member_status(self.h, st))
        return [int(x) for x in st]

    def run(self, maxiter, tol, strict=True):
        """-> [(iters, trace[2, iters+2]), ...] per member.  A failing member (K_n not positive definite, non-finite LML)
        stops alone; the others run to completion.  strict=True raises afterwards like the reference's failing
        CausalStump() call would; strict=False returns everything and leaves the verdict to member_status()."""
        its = (C.c_int * self.B)()
        traces = np.zeros((self.B, maxiter + 2, 2))
        code = self.lib.dll.cs_fit_run(self.h, int(maxiter), float(tol), its, _ptr(traces))
        if strict or (code < 0 and code != CS_ERR_NONFINITE):
            self.lib.check(code)
        return [(int(its[b]), traces[b, : int(its[b]) + 2].T.copy()) for b in range(self.B)]


def run_many(fits, maxiter, tol, strict=True):
    """cs_fit_run_many: drive several handles (Fit or FitBatch) concurrently; returns [(iters, trace), ...]
    with one entry per fit, batched handles expanded member by member in order.  strict=False: a failing member does
    not raise (see FitBatch.run); query member_status() of the handles."""
    lib = fits[0].lib
    n = len(fits)
    sizes = [getattr(f, "B", 1) for f in fits]
    total = sum(sizes)
    hs = (C.c_void_p * n)(*[f.h for f in fits])
    iters = (C.c_int * total)()
    traces = [np.zeros((b, maxiter + 2, 2)) for b in sizes]
    tp = (dp * n)(*[_ptr(t) for t in traces])
    code = lib.dll.cs_fit_run_many(hs, n, int(maxiter), float(tol), iters, tp)
    if strict or (code < 0 and code != CS_ERR_NONFINITE):
        lib.check(code)
    out, k = [], 0
    for t, b in zip(traces, sizes):
        for m in range(b):
            it = int(iters[k]); k += 1
            out.append((it, t[m, : it + 2].T.copy()))
    return out


_default = None


def load() -> CsLib:
    """The product library; raises if it is not built or no CUDA device is present."""
    global _default
    if _default is None:
        _default = CsLib(DEFAULT_LIB)
    return _default
