"""ctypes binding of oracle/_ref/libcausalstump_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's OWN C++ (src/kernel_GP_SE_cpp.cpp, src/kernel_TP_SE_cpp.cpp,
src/optimizer_cpp.cpp of mazphilip/CausalStump, compiled unmodified by oracle/Makefile
against oracle/shim/RcppArmadillo.h) behind the flat driver oracle/ref_driver.cpp.  It is
used to pin oracle/causalstump_oracle.py (tests/test_oracle_vs_ref.py) and to generate
the golden vectors of tests/golden/golden_ref_v1.npz (tests/golden/make_golden_ref.py).
It is built only where /root/reference exists (this container); the GPU box uses the
prebuilt .so that travels with the snapshot, or the committed golden vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libcausalstump_ref.so")
REF_SRC = "/root/reference/src"
dp = C.POINTER(C.c_double)


def build(force=False):
    """make -C oracle (needs /root/reference); returns the path or None when unavailable."""
    if os.path.isdir(REF_SRC):
        cmd = ["make", "-C", HERE] + (["-B"] if force else [])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB if os.path.exists(LIB) else None


def available():
    return os.path.exists(LIB)


def _f(a):
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=["F", "A"])


def _p(a):
    return a.ctypes.data_as(dp)


def build_glue(c_abi_lib, out):
    """Compiles the R-side glue of the drop-in (causalstump_b200/r/src/b200_glue.cpp) against oracle/shim/Rcpp.h and links
    it to `c_abi_lib`, a library exporting the C ABI of include/cs_b200.h (the SIMT-emulated one on CPU, the CUDA one on
    the GPU box).  Returns `out`."""
    root = os.path.dirname(HERE)
    d, f = os.path.split(os.path.abspath(c_abi_lib))
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(HERE, "shim"),
           "-I" + os.path.join(root, "include"), os.path.join(root, "causalstump_b200", "r", "src", "b200_glue.cpp"),
           os.path.join(HERE, "glue_driver.cpp"), "-o", out, "-L" + d, "-l:" + f, "-Wl,-rpath," + d]
    subprocess.check_call(cmd)
    return out


def host_lapack():
    """(ctypes library, description) of an LP64 LAPACK in this process: SciPy's bundled OpenBLAS (`scipy_dpotrf_`, ...)."""
    import scipy.linalg  # noqa: F401  (loads the library)
    from threadpoolctl import threadpool_info
    for info in threadpool_info():
        if info.get("user_api") != "blas":
            continue
        try:
            dll = C.CDLL(info["filepath"])
            for pre, suf in (("scipy_", "_"), ("", "_")):
                if all(hasattr(dll, pre + f + suf) for f in ("dpotrf", "dpotri", "dgetrf")):
                    return dll, pre, suf, "%s %s (%s)" % (info.get("internal_api"), info.get("version"), info.get("architecture"))
        except OSError:
            continue
    return None


class RefLib:
    """prefix "ref": the reference's own compiled routines; prefix "glue": the drop-in's R glue -> C ABI -> kernels."""

    def __init__(self, path=LIB, prefix="ref"):
        self.dll = C.CDLL(path)
        self.prefix = prefix
        for name in ("kernmat", "invkernel", "mu_solution", "grad", "stats", "opt_step"):
            fn = getattr(self.dll, prefix + "_" + name)
            fn.restype = C.c_int
            setattr(self.dll, "ref_" + name, fn)  # the methods below call ref_*
        self.lapack = None

    def use_host_lapack(self, on=True):
        """inv_sympd / log_det through the host LAPACK (what Armadillo links) instead of the shim's plain loops."""
        if not on:
            self.dll.ref_set_lapack(None, None, None)
            self.lapack = None
            return None
        found = host_lapack()
        if found is None:
            raise RuntimeError("no LP64 LAPACK (dpotrf_/dpotri_/dgetrf_) found in this process")
        dll, pre, suf, desc = found
        self._lapack_dll = dll
        ptr = [C.cast(getattr(dll, pre + f + suf), C.c_void_p) for f in ("dpotrf", "dpotri", "dgetrf")]
        self.dll.ref_set_lapack.argtypes = [C.c_void_p] * 3
        self.dll.ref_set_lapack.restype = None
        self.dll.ref_set_lapack(*ptr)
        self.lapack = desc
        return desc

    def log_det(self, A):
        A = _f(A)
        val, sign = C.c_double(0.0), C.c_double(0.0)
        rc = self.dll.ref_log_det(A.shape[0], _p(A), C.byref(val), C.byref(sign))
        assert rc == 0
        return val.value, sign.value

    def kernmat(self, kind, X1, X2, Z1, Z2, Lm, La, lambdam, lambda_a):
        X1, X2, Z1, Z2, Lm, La = map(_f, (X1, X2, Z1, Z2, Lm, La))
        n1, p = X1.shape
        n2 = X2.shape[0]
        K, Km, Ka = (np.empty((n1, n2), order="F") for _ in range(3))
        rc = self.dll.ref_kernmat(C.c_int(kind), n1, n2, p, _p(X1), _p(X2), _p(Z1), _p(Z2), _p(Lm), _p(La),
                                  C.c_double(lambdam), C.c_double(lambda_a), _p(K), _p(Km), _p(Ka))
        assert rc == 0
        return K, Km, Ka

    def invkernel(self, z, w, K, sigma):
        z, w, K = map(_f, (z, w, K))
        n = K.shape[0]
        out = np.empty((n, n), order="F")
        rc = self.dll.ref_invkernel(n, _p(z), _p(w), _p(K), C.c_double(sigma), _p(out))
        if rc != 0:
            raise np.linalg.LinAlgError("inv_sympd(): matrix is singular or not positive definite")
        return out

    def mu_solution(self, y, z, invK):
        y, z, invK = map(_f, (y, z, invK))
        mu = C.c_double(0.0)
        self.dll.ref_mu_solution(len(y), _p(y), _p(z), _p(invK), C.byref(mu))
        return mu.value

    def grad(self, kind, y, X, z, w, K, Km, Ka, invK, params):
        y, X, z, w, K, Km, Ka, invK, params = map(_f, (y, X, z, w, K, Km, Ka, invK, params))
        n, p = X.shape
        g = np.empty(len(params))
        st = np.empty(2)
        rc = self.dll.ref_grad(C.c_int(kind), n, p, _p(y), _p(X), _p(z), _p(w), _p(K), _p(Km), _p(Ka), _p(invK), _p(params),
                               _p(g), _p(st))
        assert rc == 0
        return g, st

    def stats(self, kind, p, y, K, invK, params):
        y, K, invK, params = map(_f, (y, K, invK, params))
        st = np.empty(2)
        rc = self.dll.ref_stats(C.c_int(kind), len(y), p, _p(y), _p(K), _p(invK), _p(params), _p(st))
        assert rc == 0
        return st

    def opt_step(self, optim, kind, p, it, lr, beta1, beta2, eps, m, v, grad, para):
        m, v, grad, para = (np.array(_f(a), copy=True) for a in (m, v, grad, para))
        rc = self.dll.ref_opt_step(C.c_int(optim), C.c_int(kind), p, C.c_double(it), C.c_double(lr), C.c_double(beta1),
                                   C.c_double(beta2), C.c_double(eps), _p(m), _p(v), _p(grad), _p(para))
        assert rc == 0
        return m, v, para

    # one full evaluation the way para_update chains the routines (R/kernel_GP_SE_class.R:45-60)
    def eval_lml_grad(self, kind, y, X, z, w, params):
        X = _f(X)
        p = X.shape[1]
        params = _f(params)
        off = 3 if kind == 0 else 4
        Lm, La = params[off:off + p], params[off + p:off + 2 * p]
        K, Km, Ka = self.kernmat(kind, X, X, z, z, Lm, La, params[1], params[2] if kind == 0 else params[3])
        invK = self.invkernel(z, w, K, params[0])
        g, st = self.grad(kind, y, X, z, w, K, Km, Ka, invK, params)
        mu = self.mu_solution(y, z, invK)
        return g, st, mu, dict(K=K, Km=Km, Ka=Ka, invK=invK)
