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


# --------------------------------------------------------------------------------------------------------------
# Host mirror of the DEVICE generator (csrc/sim.cuh, cs_sim_run): the same Philox-4x32-10 counters, so a replicate
# produced on the GPU can be re-created here draw for draw.  Discrete draws (beta categories, Bernoulli covariates,
# treatment, the train / test split) are identical; normals and sums agree to rounding (libm vs CUDA exp / log / cos,
# reduction order).  Input generation only -- nothing here is on the measured path.
# --------------------------------------------------------------------------------------------------------------
_S_BETA, _S_XNORM, _S_XBIN, _S_Z, _S_N0, _S_N1, _S_SPLIT, _S_RESTART = range(8)
_SIM_TAG = 0x48494C4C


def philox4x32(c0, c1, c2, c3, seed):
    """Philox-4x32-10 (csrc/predict.cuh: csp::Philox::block) on uint32 arrays; key = (seed & 0xffffffff, seed >> 32)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [h1 ^ c[1] ^ k0, l1, h0 ^ c[3] ^ k1, l0]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def _u01(hi, lo):
    bits = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (bits.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _sim_uniform(index, stream, rep, seed):
    o = philox4x32(index, stream, rep, _SIM_TAG, seed)
    return _u01(o[0], o[1])


def _sim_normal(index, stream, rep, seed):
    o = philox4x32(index, stream, rep, _SIM_TAG, seed)
    u1, u2 = _u01(o[0], o[1]), _u01(o[2], o[3])
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586476925 * u2)


def hill_unit_philox(replicate, restart, n=747, p=25, X=None, z=None, test_frac=0.1, seed=565):
    """One (replicate, restart) unit exactly as sim_generate_kernel builds it: response surface B
    (test/sec5-3_Hill2011.R:49-66), the 10 % test split (:171), norm_variables on the training rows (R/utilities.R:1-33)
    and parainit (R/kernel_GP_SE_class.R:10-20).  Returns raw and normalised pieces plus the initial parameter vector."""
    import math
    if X is None:
        ncont = min(6, p)
        idx = np.arange(n * p, dtype=np.uint64)
        j = (idx // np.uint64(n)).astype(np.int64)
        xn = _sim_normal(idx, _S_XNORM, replicate, seed)
        means = np.array(_BIN_MEANS)[np.maximum(j - ncont, 0) % 19]
        xb = (_sim_uniform(idx, _S_XBIN, replicate, seed) < means).astype(np.float64) + ((j - ncont) == 7)
        X = np.where(j < ncont, xn, xb).reshape(p, n).T.copy()
        z = (_sim_uniform(np.arange(n, dtype=np.uint64), _S_Z, replicate, seed) < 0.186).astype(np.float64)
        if z.sum() == 0:
            z[0] = 1.0
    else:
        X = np.array(X, dtype=np.float64)
        z = np.array(z, dtype=np.float64)
        n, p = X.shape
    u = _sim_uniform(np.arange(p + 1, dtype=np.uint64), _S_BETA, replicate, seed)
    beta = np.select([u < 0.6, u < 0.7, u < 0.8, u < 0.9], [0.0, 0.1, 0.2, 0.3], 0.4)
    lin = np.full(n, beta[0])
    for jj in range(p):
        lin = (X[:, jj] + 0.5) * beta[jj + 1] + lin        # two roundings per term, like the device (no FMA)
    y0 = np.exp(lin)
    offset = float(np.mean((lin - y0)[z == 1])) - 4.0
    y1 = lin - offset
    ii = np.arange(n, dtype=np.uint64)
    e0, e1 = _sim_normal(ii, _S_N0, replicate, seed), _sim_normal(ii, _S_N1, replicate, seed)
    y = np.where(z == 1, y1 + e1, y0 + e0)
    key = _sim_uniform(ii, _S_SPLIT, replicate, seed)
    nte = int(math.ceil(n * test_frac))
    order = np.lexsort((np.arange(n), key))
    is_test = np.zeros(n, dtype=bool)
    is_test[order[:nte]] = True
    train, test = np.flatnonzero(~is_test), np.flatnonzero(is_test)
    from .main import norm_variables
    nr = norm_variables(y[train], X[train])
    mom = nr["moments"]
    Xall = norm_variables(X=X, moments=mom)["X"]
    yn = nr["y"]
    ab = _sim_uniform(np.array([2 * restart, 2 * restart + 1], dtype=np.uint64), _S_RESTART, replicate, seed)
    params = np.concatenate([[math.log(float(np.var(yn, ddof=1))), math.log(0.2 + 0.8 * ab[0]), math.log(0.2 + 0.8 * ab[1])],
                             np.full(p, -0.1), np.full(p, 0.1), [float(np.mean(yn))]])
    return dict(y=y, X=X, z=z, tau=y1 - y0, y1=y1, y0=y0, is_test=is_test, train=train, test=test, moments=mom,
                Xtrain=np.asfortranarray(nr["X"]), ytrain=yn, ztrain=z[train], Xall=Xall, params=params, beta=beta)
