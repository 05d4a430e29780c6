"""Host-side mirrors of the reference's 11 registered `.Call` routines
(/root/reference/R/RcppExports.R:4-46, src/RcppExports.cpp:187-200) -- same names,
argument order and list-shaped parameters -- bound to the C ABI of
libcausalstump_b200.so.  R lists are Python dicts in the reference's list order.
There is no CPU arithmetic here: every function is a device call and raises if the
CUDA library or a device is missing.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

from . import _lib

GP_ORDER = ("sigma", "lambdam", "lambdaa", "Lm", "La", "mu")
TP_ORDER = ("sigma", "lambdam", "nu", "lambda0", "Lm", "La", "mu")

_active = None


def use_library(lib):
    """Bind the mirrors to a specific loaded CsLib (tests use this); default = product lib."""
    global _active
    _active = lib


def lib():
    return _active if _active is not None else _lib.load()


def _mat(X):
    X = np.asarray(X, dtype=np.float64)
    return X[:, None] if X.ndim == 1 else X


def pack(par):
    out = []
    for v in par.values():
        out.extend(np.atleast_1d(np.asarray(v, dtype=np.float64)).tolist())
    return np.array(out, dtype=np.float64)


def unpack(vec, template):
    """flat vector -> dict shaped like `template` (scalars stay scalars)."""
    out, k = OrderedDict(), 0
    for name, v in template.items():
        if np.ndim(v) == 0:
            out[name] = float(vec[k]); k += 1
        else:
            m = len(v)
            out[name] = np.array(vec[k:k + m], dtype=np.float64); k += m
    return out


def _kind(par):
    return _lib.CS_TP if "lambda0" in par else _lib.CS_GP


def kernmat_GP_SE_cpp(X1, X2, Z1, Z2, para):
    o = lib().kernmat(_mat(X1), _mat(X2), Z1, Z2, para["Lm"], para["La"], para["lambdam"], para["lambdaa"])
    return OrderedDict(Kmat=o["K"], Km=o["Km"], Ka=o["Ka"])


def kernmat_TP_SE_cpp(X1, X2, Z1, Z2, para):
    o = lib().kernmat(_mat(X1), _mat(X2), Z1, Z2, para["Lm"], para["La"], para["lambdam"], para["lambda0"])
    return OrderedDict(Smat=o["K"], Sm=o["Km"], Sa=o["Ka"])


def invkernel_cpp(z, w, mymat, parameters):
    return lib().invkernel(w, mymat, parameters["sigma"])


def mu_solution_cpp(y, z, invKmat, parameters):
    return lib().mu_solution(y, invKmat)


def _grad(kind, y, X, z, w, Kmat, Km, Ka, invK, parameters):
    g, st = lib().grad(kind, y, _mat(X), z, w, pack(parameters), invK=invK, Kmat=Kmat, Km=Km, Ka=Ka)
    return OrderedDict(gradients=unpack(g, parameters), stats=st)


def grad_GP_SE_cpp(y, X, z, w, Kmat, Km, Ka, invKmatn, parameters):
    return _grad(_lib.CS_GP, y, X, z, w, Kmat, Km, Ka, invKmatn, parameters)


def grad_TP_SE_cpp(y, X, z, w, Smat, Sm, Sa, invSmatn, parameters):
    return _grad(_lib.CS_TP, y, X, z, w, Smat, Sm, Sa, invSmatn, parameters)


def stats_GP_SE(y, Kmat, invKmatn, parameters):
    return lib().stats(_lib.CS_GP, y, Kmat, invKmatn, parameters["mu"])


def stats_TP_SE(y, Kmat, invKmatn, parameters):
    return lib().stats(_lib.CS_TP, y, Kmat, invKmatn, parameters["mu"], math.exp(parameters["nu"]))


def _positional(state, grad, para):
    """The reference loops over the first len(state) list entries positionally."""
    k = len(state)
    names = list(para.keys())[:k]
    sub = OrderedDict((nm, para[nm]) for nm in names)
    gsub = OrderedDict((nm, list(grad.values())[i]) for i, nm in enumerate(names))
    return names, sub, gsub


def Nesterov_cpp(learn_rate, momentum, nu, grad, para):
    names, sub, gsub = _positional(nu, grad, para)
    m2, _, p2 = lib().opt_step(_lib.CS_OPT_NESTEROV, 1.0, learn_rate, momentum, 0.0, 0.0, pack(nu), None, pack(gsub),
                               pack(sub))
    out_para = OrderedDict(para)
    out_para.update(unpack(p2, sub))
    return OrderedDict(nu=unpack(m2, nu), parameters=out_para)


def _adamlike(code, it, learn_rate, beta1, beta2, eps, m, v, grad, para):
    names, sub, gsub = _positional(m, grad, para)
    m2, v2, p2 = lib().opt_step(code, it, learn_rate, beta1, beta2, eps, pack(m), pack(v), pack(gsub), pack(sub))
    out_para = OrderedDict(para)
    out_para.update(unpack(p2, sub))
    return OrderedDict(m=unpack(m2, m), v=unpack(v2, v), parameters=out_para)


def Nadam_cpp(iter, learn_rate, beta1, beta2, eps, m, v, grad, para):  # noqa: A002 (reference name)
    return _adamlike(_lib.CS_OPT_NADAM, iter, learn_rate, beta1, beta2, eps, m, v, grad, para)


def Adam_cpp(iter, learn_rate, beta1, beta2, eps, m, v, grad, para):  # noqa: A002
    return _adamlike(_lib.CS_OPT_ADAM, iter, learn_rate, beta1, beta2, eps, m, v, grad, para)
