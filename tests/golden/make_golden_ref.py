"""Regenerates tests/golden/golden_ref_v1.npz from the reference's OWN C++.

Every number in the fixture is produced by oracle/_ref/libcausalstump_ref.so, i.e. by
/root/reference/src/{kernel_GP_SE_cpp,kernel_TP_SE_cpp,optimizer_cpp}.cpp compiled unmodified
(oracle/Makefile, against oracle/shim/RcppArmadillo.h) -- NOT by the NumPy oracle.  The only
restated part is the R glue that chains the routines (R/kernel_GP_SE_class.R:45-73,
R/kernel_TP_SE_class.R:45-70, R/main.R:79-88, R/optimizer_classes.R), written out below with
the reference line numbers; R itself is not installed in this image.

Inputs are seeded NumPy data (causalstump_b200/hill.py), normalised like R/utilities.R.
Consumers: tests/test_golden_ref.py (oracle, CPU) and tests/test_gpu_parity.py (CUDA path, B200).

    python tests/golden/make_golden_ref.py      # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from causalstump_b200 import hill  # noqa: E402
from causalstump_b200.main import norm_variables  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ref_v1.npz")

CASES = [  # name, kind, n, p, readme, weights, optimizer, iterations, lr, momentum
    ("gp_readme", 0, 120, 1, True, False, "Nadam", 8, 0.02, 0.0),
    ("gp_hill25", 0, 180, 25, False, True, "Adam", 6, 0.01, 0.0),
    ("tp_hill", 1, 130, 6, False, False, "Nadam", 6, 0.02, 0.0),
    ("gp_nesterov", 0, 90, 3, False, False, "Nesterov", 5, 0.001, 0.4),
    ("tp_hill25", 1, 160, 25, False, True, "Adam", 4, 0.01, 0.0),
]
OPT = {"Adam": 0, "Nadam": 1, "Nesterov": 2}


def init_params(kind, p, y, draws, nu=50.0):
    """parainit (R/kernel_GP_SE_class.R:10-20, R/kernel_TP_SE_class.R:11-22) with the two runif draws as inputs."""
    sig = np.log(np.var(y, ddof=1))
    if kind == 0:
        head = [sig, np.log(draws[0]), np.log(draws[1])]
    else:
        head = [sig, np.log(draws[0]), np.log(nu), np.log(draws[1])]
    return np.array(head + [-0.1] * p + [0.1] * p + [np.mean(y)])


def lists_of(kind, p, params):
    off = 3 if kind == 0 else 4
    return params[off:off + p], params[off + p:off + 2 * p], params[1], params[2] if kind == 0 else params[3]


def getinv_kernel(ref, kind, p, X, z, w, params):
    """R/kernel_GP_SE_class.R:31-44."""
    Lm, La, lm, la = lists_of(kind, p, params)
    K, Km, Ka = ref.kernmat(kind, X, X, z, z, Lm, La, lm, la)
    return K, Km, Ka, ref.invkernel(z, w, K, params[0])


def fit_loop(ref, kind, p, y, X, z, w, params, optim, iters, lr, momentum):
    """R/main.R:73-88 with para_update (R/kernel_GP_SE_class.R:45-60 / R/kernel_TP_SE_class.R:45-59); tol = 0."""
    params = params.copy()
    m = np.zeros_like(params)  # list2zero(parameters): includes mu (R/optimizer_classes.R:29-33)
    v = np.zeros_like(params)
    stats = np.zeros((2, iters + 2))
    for it in range(1, iters + 1):
        K, Km, Ka, invK = getinv_kernel(ref, kind, p, X, z, w, params)
        g, st = ref.grad(kind, y, X, z, w, K, Km, Ka, invK, params)
        b1 = momentum if optim == "Nesterov" else 0.9
        m, v, params = ref.opt_step(OPT[optim], kind, p, float(it), lr, b1, 0.999, 1e-8, m, v, g, params)
        params[-1] = ref.mu_solution(y, z, invK)  # mean_solution with the PRE-update inverse (:68-73)
        stats[:, it] = st
    K, Km, Ka, invK = getinv_kernel(ref, kind, p, X, z, w, params)  # get_train_stats (R/main.R:88)
    stats[:, iters + 1] = ref.stats(kind, p, y, K, invK, params)
    return params, stats, invK


def predict(ref, kind, p, y, X, z, X2, z2, params, invK):
    """KernelClass$predict (R/kernel_GP_SE_class.R:74-85): map and diag(cov)."""
    Lm, La, lm, la = lists_of(kind, p, params)
    KxX = ref.kernmat(kind, X2, X, z2, z, Lm, La, lm, la)[0]
    Kxx = ref.kernmat(kind, X2, X2, z2, z2, Lm, La, lm, la)[0]
    mu = params[-1]
    mp = KxX @ invK @ (y - mu) + mu
    cov = Kxx - KxX @ invK @ KxX.T + np.exp(params[0]) * np.eye(len(z2))
    return mp, np.diag(cov).copy()


def predict_treat(ref, kind, p, y, X, z, X2, params, invK):
    """KernelClass$predict_treat (R/kernel_GP_SE_class.R:86-105; the TP class adds 1e-6 to diag(cov),
    R/kernel_TP_SE_class.R:110)."""
    Lm, La, lm, la = lists_of(kind, p, params)
    z2 = np.ones(X2.shape[0])
    KxX = ref.kernmat(kind, X2, X, z2, z, Lm, La, lm, la)[2]
    Kxx = ref.kernmat(kind, X2, X2, z2, z2, Lm, La, lm, la)[2]
    mp = KxX @ invK @ (y - params[-1])
    cov = Kxx - KxX @ invK @ KxX.T
    if kind == 1:
        cov = cov + 1e-6 * np.eye(X2.shape[0])
    return mp, np.diag(cov).copy(), float(np.mean(mp)), float(np.sum(cov) / len(y) ** 2)


def build():
    if rb.build() is None:
        raise SystemExit("oracle/_ref cannot be built here (no /root/reference)")
    ref = rb.RefLib()
    out = {}
    for name, kind, n, p, readme, weights, optim, iters, lr, mom in CASES:
        d = hill.readme_example(n) if readme else hill.hill_surface_b(n=n, p=p, replicate=11)
        nr = norm_variables(d["y"], d["X"])
        y, X, z = np.ascontiguousarray(nr["y"]), np.asfortranarray(nr["X"]), np.ascontiguousarray(d["z"], dtype=np.float64)
        p = X.shape[1]
        w = np.random.default_rng(5).uniform(0.5, 2.0, n) if weights else np.ones(n)
        par0 = init_params(kind, p, y, (0.45, 0.8) if kind == 0 else (0.45, 0.35))
        g, st, mu, mats = ref.eval_lml_grad(kind, y, X, z, w, par0)
        parK, trace, invK = fit_loop(ref, kind, p, y, X, z, w, par0, optim, iters, lr, mom)
        X2 = np.asfortranarray(X[: min(n, 29)] * 0.95 + 0.01)
        z2 = 1.0 - z[: X2.shape[0]]
        pm, pv = predict(ref, kind, p, y, X, z, X2, z2, parK, invK)
        tm, tv, ate, atev = predict_treat(ref, kind, p, y, X, z, X2, parK, invK)
        pre = name + "/"
        out.update({
            pre + "meta": np.array([kind, n, p, iters, OPT[optim], lr, mom]),
            pre + "y": y, pre + "X": X, pre + "z": z, pre + "w": w, pre + "init_params": par0,
            pre + "grad": g, pre + "stats": st, pre + "mu_sol": np.array(mu),
            pre + "invK_diag": np.diag(mats["invK"]).copy(), pre + "invK_row0": mats["invK"][0].copy(),
            pre + "K_row0": mats["K"][0].copy(), pre + "Km_row0": mats["Km"][0].copy(), pre + "Ka_row0": mats["Ka"][0].copy(),
            pre + "fit_params": parK, pre + "fit_trace": trace,
            pre + "X2": X2, pre + "z2": z2, pre + "pred_map": pm, pre + "pred_var": pv,
            pre + "treat_map": tm, pre + "treat_var": tv, pre + "treat_ate": np.array(ate), pre + "treat_ate_var": np.array(atev),
        })
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    build()
