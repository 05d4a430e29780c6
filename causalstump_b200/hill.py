"""Synthetic inputs shaped like the reference's workloads.

* ``hill_surface_b``  -- Hill (2011) IHDP response surface B as drawn in
  /root/reference/test/sec5-3_Hill2011.R:49-66, on synthetic covariates with
  the IHDP marginals (6 standardised continuous + 19 binary, `first` coded
  {1,2}); treatment ~ Bernoulli(0.186) (139/747 in the reference workspace).
* ``readme_example``  -- the n=120, p=1 toy of /root/reference/README.md:23-46.

R's Mersenne-Twister stream is not available, so seeds are NumPy
``default_rng`` seeds (replicate i uses 565 + 5*i, mirroring :159).  This is
input generation only; nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np

# Bernoulli means of the 19 binary IHDP covariates (sex ... was); index 7 is `first` (+1 -> {1,2})
_BIN_MEANS = (.51, .09, .52, .36, .27, .22, .36, .46, .14, .96, .59, .96, .14, .14, .16, .08, .07, .13, .16)


def hill_covariates(n, rng, p=25):
    ncont = min(6, p)
    X = np.empty((n, p), dtype=np.float64)
    X[:, :ncont] = rng.standard_normal((n, ncont))
    for j in range(ncont, p):
        m = _BIN_MEANS[(j - ncont) % len(_BIN_MEANS)]
        X[:, j] = (rng.random(n) < m).astype(np.float64)
        if (j - ncont) == 7:
            X[:, j] += 1.0
    return X


def hill_surface_b(n=747, p=25, replicate=0, seed=None, X=None, z=None):
    """Returns dict(y, X, z, tau, y1, y0) -- raw (un-normalised) data.  With X (and z) given -- e.g. the real
    IHDP covariates and treatment indicator read by causalstump_b200.rdata from the reference's workspace
    files -- only the response surface is drawn, as test/sec5-3_Hill2011.R:49-66 does."""
    rng = np.random.default_rng(565 + 5 * replicate if seed is None else seed)
    if X is None:
        X = hill_covariates(n, rng, p)
    else:
        X = np.array(X, dtype=np.float64)
        n, p = X.shape
    if z is None:
        z = (rng.random(n) < 0.186).astype(np.float64)
    else:
        z = np.asarray(z, dtype=np.float64)
    if z.sum() == 0:
        z[0] = 1.0
    beta = rng.choice(np.array([0.0, 0.1, 0.2, 0.3, 0.4]), size=p + 1, replace=True, p=[.6, .1, .1, .1, .1])
    D = np.column_stack([np.ones(n), X + 0.5])
    lin = D @ beta
    y0hat = np.exp(lin)
    y1hat = lin.copy()
    offset = float(np.mean(y1hat[z == 1] - y0hat[z == 1])) - 4.0
    y1hat = lin - offset
    Y0 = rng.normal(y0hat, 1.0)
    Y1 = rng.normal(y1hat, 1.0)
    y = np.where(z == 1, Y1, Y0)
    return dict(y=y, X=X, z=z, tau=y1hat - y0hat, y1=y1hat, y0=y0hat)


def readme_example(n=120, seed=1231):
    """README.md:23-46 toy: p=1, overlap-limited covariate, clipped surfaces."""
    rng = np.random.default_rng(seed)
    Z = (rng.random(n) < 0.3).astype(np.float64)
    n1 = int(Z.sum())
    X1 = rng.normal(40.0, 10.0, n1)
    X0 = rng.normal(20.0, 10.0, n - n1)
    X0[X0 < 0] = 0.01
    X = np.empty(n)
    X[Z == 1] = X1
    X[Z == 0] = X0
    y0 = np.clip(72.0 + 3.0 * np.sqrt(np.maximum(X, 0.0)), 60.0, 120.0)
    y1 = np.clip(90.0 + np.exp(0.06 * X), 60.0, 120.0)
    Y0 = rng.normal(y0, 1.0)
    Y1 = rng.normal(y1, 1.0)
    Y = Y0 * (1 - Z) + Y1 * Z
    return dict(y=Y, X=X[:, None], z=Z, tau=y1 - y0, y1=y1, y0=y0)


def restart_draws(replicate, restart, prior=False):
    """The two runif draws of parainit (R/kernel_GP_SE_class.R:15-16,
    R/kernel_TP_SE_class.R:16,18) for (replicate, restart)."""
    rng = np.random.default_rng(10_000 * restart + replicate)
    a = rng.uniform(0.2, 1.0)
    b = rng.uniform(0.1, 0.7) if prior else rng.uniform(0.2, 1.0)
    return float(a), float(b)
