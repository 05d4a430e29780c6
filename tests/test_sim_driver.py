"""Simulation driver on the device (cs_sim_run / cs_sim_generate, csrc/sim.cuh) against its host restatement:
generator = causalstump_b200.hill.hill_unit_philox (same Philox counters), fit / predict / treatment = the oracle,
metric block = replicates.update_results (test/sec5-3_Hill2011.R:104-152).  CPU: the kernels compiled for the SIMT
emulator at small n; GPU (-m gpu): the CUDA library at the reference's size (n = 747, p = 25, real loop settings)."""
import math

import numpy as np
import pytest

from causalstump_b200 import _lib, hill, replicates as rp
from oracle import causalstump_oracle as o


def host_unit(replicate, restart, n, p, maxiter, tol, lr, seed, X=None, z=None):
    h = hill.hill_unit_philox(replicate, restart, n=n, p=p, X=X, z=z, seed=seed)
    K = o.KernelClass_GP_SE(p, np.ones(len(h["train"])))
    K.parameters = o.unpack_params(h["params"], "GP", p)
    opt = o.optNadam(lr, 0.9, 0.999)
    opt.initOpt(K)
    y, Xt, zt = h["ytrain"], h["Xtrain"], h["ztrain"]
    prev, iters = 0.0, maxiter
    for it in range(1, maxiter + 1):                      # R/main.R:79-84
        st = K.para_update(it, y, Xt, zt, opt, faithful=False)
        change = abs(st[1] - prev)
        prev = st[1]
        if change < tol and it > 3:
            iters = it
            break
    lml = K.get_train_stats(y, Xt, zt)[1]
    mom = h["moments"]
    sd, vy = math.sqrt(mom["varY"]), mom["varY"]
    pred = K.predict(y, Xt, zt, h["Xall"], h["z"])["map"] * sd + mom["meanY"]
    nn = len(h["y"])
    tau_hat, tau_ci, ate_ci = np.empty(nn), np.empty((nn, 2)), []
    for idx in (h["train"], h["test"]):
        t = K.predict_treat(y, Xt, zt, h["Xall"][idx])
        var = np.diag(t["cov"])
        tau_hat[idx] = sd * t["map"]
        tau_ci[idx, 0] = vy * (-1.96 * var) + sd * t["map"]
        tau_ci[idx, 1] = vy * (1.96 * var) + sd * t["map"]
        half = 1.96 * math.sqrt(np.sum(t["cov"]) / len(y) ** 2)
        ate_ci.append((vy * (-half) + sd * t["ate_map"], vy * half + sd * t["ate_map"]))
    m = rp.update_results(h["test"], pred, tau_hat, tau_ci, ate_ci[0], ate_ci[1], h["y"], h["z"], h["y1"], h["y0"])
    return np.array([m[k] for k in rp.METRICS]), iters, lml


def check_generator(lib, n, p, X=None, z=None):
    for rep, rs in ((0, 0), (7, 3)):
        d = lib.sim_generate(rep, rs, n=n, p=p, X=X, z=z, seed=99)
        h = hill.hill_unit_philox(rep, rs, n=n, p=p, X=X, z=z, seed=99)
        assert np.array_equal(d["z"], h["z"]) and np.array_equal(d["is_test"].astype(bool), h["is_test"])   # discrete draws: identical
        assert np.array_equal(d["ztrain"], h["ztrain"])
        vx = h["moments"]["varX"]
        assert np.array_equal(d["moments"][p:2 * p] == 1.0, vx == 1.0) and np.allclose(d["moments"][p:2 * p], vx, rtol=1e-13, atol=0)
        binary = np.all((h["X"] == 0) | (h["X"] == 1) | (h["X"] == 2), axis=0)
        assert np.array_equal(d["Xraw"][:, binary], h["X"][:, binary])
        for k, hk in (("Xraw", "X"), ("y", "y"), ("y1", "y1"), ("y0", "y0"), ("Xtrain", "Xtrain"), ("ytrain", "ytrain"), ("Xall", "Xall"), ("params", "params")):
            assert np.max(np.abs(d[k] - h[hk])) <= 1e-12 * max(1.0, float(np.max(np.abs(h[hk])))), k
        assert abs(d["moments"][-1] - h["moments"]["varY"]) <= 1e-13 * h["moments"]["varY"]


def check_run(lib, units, n, p, maxiter, tol, rtol, X=None, z=None, batch=3):
    out = lib.sim_run(units, n=n, p=p, X=X, z=z, maxiter=maxiter, tol=tol, lr=0.01, batch=batch, seed=99)
    assert out["metrics"].shape == (len(units), 22) and (out["status"] == 0).all()
    for u, (rep, rs) in enumerate(units):
        m, iters, lml = host_unit(rep, rs, n, p, maxiter, tol, 0.01, 99, X=X, z=z)
        assert out["iters"][u] == iters, (u, out["iters"][u], iters)
        assert abs(out["lml"][u] - lml) <= 1e-8 * abs(lml)
        both_nan = np.isnan(m) & np.isnan(out["metrics"][u])
        assert np.all(both_nan | (np.abs(out["metrics"][u] - m) <= rtol * np.maximum(np.abs(m), 1e-3))), (u, out["metrics"][u], m)
    return out


def test_device_generator_matches_the_host_mirror(emulib):
    check_generator(emulib, 120, 25)
    check_generator(emulib, 97, 3)
    rng = np.random.default_rng(3)
    check_generator(emulib, 50, 4, X=rng.normal(size=(50, 4)), z=(rng.random(50) < 0.3).astype(float))


def test_device_simulation_loop_matches_the_host_pipeline(emulib):
    check_run(emulib, [(0, 0), (0, 1), (1, 0), (2, 2)], n=70, p=3, maxiter=6, tol=0.0, rtol=1e-7)


@pytest.mark.gpu
def test_device_generator_on_gpu(gpulib):
    check_generator(gpulib, 747, 25)


@pytest.mark.gpu
def test_device_simulation_loop_on_gpu(gpulib):
    """config-4 units at the reference's size and stopping rule, on synthetic and on shared covariates"""
    check_run(gpulib, [(0, 0), (0, 1), (3, 2)], n=300, p=25, maxiter=40, tol=1e-1, rtol=1e-6, batch=2)
    rng = np.random.default_rng(5)
    X = np.column_stack([rng.normal(size=(200, 3)), (rng.random((200, 2)) < 0.4).astype(float)])
    check_run(gpulib, [(1, 0), (2, 1)], n=200, p=5, maxiter=25, tol=1e-1, rtol=1e-6, X=X, z=(rng.random(200) < 0.25).astype(float))


def test_replicate_driver_on_the_device_driver(emulib):
    """replicates.hill_units_device: the Python face of cs_sim_run returns the same row shape as the host driver
    (metrics dict over METRICS, iteration count, final LML, failure flag) and restart winners can be selected from it."""
    from causalstump_b200 import rcpp_exports as rx
    rx.use_library(emulib)
    try:
        units = [(0, 0), (0, 1), (1, 0)]
        res = rp.hill_units_device(units, n=64, p=3, maxiter=5, tol=0.0, batch=2, seed=11)
        assert [(r["replicate"], r["restart"]) for r in res] == units
        for r in res:
            assert set(r["metrics"]) == set(rp.METRICS) and r["iters"] == 5 and not r["failed"] and np.isfinite(r["lml"])
            assert r["pehe_test"] == r["metrics"]["PEHE.all.test"] and 0.0 <= r["metrics"]["coverage.ite.train"] <= 1.0
        win = rp.select_winners(res)
        assert [w["replicate"] for w in win] == [0, 1] and win[0]["lml"] == max(res[0]["lml"], res[1]["lml"])
        assert 0.0 < rp.LAST_TIMING["slot_utilisation"] <= 1.0 and rp.LAST_TIMING["iterations"] == 15
    finally:
        rx.use_library(None)
