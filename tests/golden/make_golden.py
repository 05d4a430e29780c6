"""Regenerates tests/golden/golden_v1.npz from the oracle (oracle/causalstump_oracle.py).

The reference (R + RcppArmadillo) cannot run in the build image and ships no golden
vectors for this path, so these fixtures pin the ORACLE (a restatement), not the
reference itself -- "parity unpinned", see DESIGN.md.  They are regression anchors:
tests/test_oracle.py checks the oracle still reproduces them, tests/test_gpu_parity.py
checks the CUDA path against them on the B200.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from causalstump_b200 import hill  # noqa: E402
from oracle import causalstump_oracle as o  # noqa: E402

CASES = [  # name, kind, n, p, readme, weights, optimizer, iters
    ("gp_readme", "GP", 120, 1, True, False, "Nadam", 8),
    ("gp_hill", "GP", 96, 5, False, True, "Adam", 6),
    ("tp_hill", "TP", 80, 4, False, False, "Nadam", 6),
    ("gp_nesterov", "GP", 64, 3, False, False, "Nesterov", 5),
]


def build():
    out = {}
    for name, kind, n, p, readme, weights, optim, iters in CASES:
        d = hill.readme_example(n) if readme else hill.hill_surface_b(n=n, p=p, replicate=7)
        w = np.random.default_rng(3).uniform(0.5, 2.0, n) if weights else np.ones(n)
        draws = (0.45, 0.35) if kind == "TP" else (0.45, 0.8)
        nr = o.norm_variables(d["y"], d["X"])
        y, X, z = nr["y"], nr["X"], d["z"]
        pp = X.shape[1]
        if kind == "GP":
            K = o.KernelClass_GP_SE(pp, w); K.parainit(y, *draws)
        else:
            K = o.KernelClass_TP_SE(pp, w); K.parainit(y, 50.0, *draws)
        g, st, mu = o.eval_lml_grad(kind, y, X, z, w, K.parameters)
        fit = o.CausalStump(d["y"], d["X"], d["z"], w=w, myoptim=optim, maxiter=iters, tol=1e-14, prior=(kind == "TP"),
                            nu=50.0, learning_rate=0.02, momentum=0.3, init_draws=draws, faithful=True)
        X2 = X[: min(n, 33)] + 0.01
        z2 = 1.0 - z[: X2.shape[0]]
        pr = fit.Kernel.predict(y, X, z, X2, z2)
        tr = fit.Kernel.predict_treat(y, X, z, X2)
        pre = name + "/"
        out.update({
            pre + "y_raw": d["y"], pre + "X_raw": d["X"], pre + "z": z, pre + "w": w, pre + "tau": d["tau"],
            pre + "y": y, pre + "X": X, pre + "init_params": o.pack_params(K.parameters),
            pre + "grad": o.pack_params(g), pre + "stats": st, pre + "mu_sol": np.array(mu),
            pre + "fit_params": o.pack_params(fit.Kernel.parameters), pre + "fit_trace": fit.stats,
            pre + "X2": X2, pre + "z2": z2, pre + "pred_map": pr["map"], pre + "pred_var": np.diag(pr["cov"]),
            pre + "treat_map": tr["map"], pre + "treat_var": np.diag(tr["cov"]),
            pre + "treat_ate": np.array(tr["ate_map"]), pre + "treat_ate_var": np.array(np.sum(tr["cov"]) / n ** 2),
            pre + "meta": np.array([0 if kind == "GP" else 1, n, pp, iters,
                                    {"Adam": 0, "Nadam": 1, "Nesterov": 2}[optim]], dtype=np.float64),
        })
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **build())
    print(path, os.path.getsize(path), "bytes")
