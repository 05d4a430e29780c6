"""Replicate / restart sharding across GPUs (config 4 of BASELINE.json; mirrors the
reference's `parSapplyLB(cl, 1:100, run)` over a FORK cluster,
/root/reference/test/sec5-3_Hill2011.R:611-616).

Units are independent (replicate, restart) fits: one process per GPU, static round-robin
assignment, NO collective on the data path; results (a few floats per unit) are gathered
on rank 0 with `all_gather_object` and the restart with the highest final log marginal
likelihood wins per replicate.
"""
from __future__ import annotations

import math

import numpy as np


def unit_list(n_replicates, n_restarts):
    return [(rep, rs) for rep in range(n_replicates) for rs in range(n_restarts)]


def shard(units, rank, world):
    """Static round-robin: rank r takes units r, r+world, ...  Every unit exactly once."""
    return list(units[rank::world])


def hill_unit(replicate, restart, n=747, p=25, test_frac=0.1, maxiter=3000, tol=1e-1, lr=0.01, prior=False, nu=2000):
    """One replicate fit of test/sec5-3_Hill2011.R:157-365 (GP arm): draw surface B, 10 % test
    split, fit on the training split, predict on all units, treatment on train and test."""
    import causalstump_b200 as cb
    from . import hill
    d = hill.hill_surface_b(n=n, p=p, replicate=replicate)
    rng = np.random.default_rng(565 + 5 * replicate)
    ntest = int(math.ceil(n * test_frac))
    test = np.sort(rng.choice(n, ntest, replace=False))
    train = np.setdiff1d(np.arange(n), test)
    fit = cb.CausalStump(d["y"][train], d["X"][train], d["z"][train], maxiter=maxiter, tol=tol, learning_rate=lr,
                         myoptim="Nadam", prior=prior, nu=nu, init_draws=hill.restart_draws(replicate, restart, prior),
                         verbose=False)
    pred = cb.predict(fit, X=d["X"], z=d["z"])
    t_tr = cb.treatment(fit, X=d["X"][train])
    t_te = cb.treatment(fit, X=d["X"][test])
    tau = d["tau"]
    return dict(replicate=replicate, restart=restart, iters=int(fit.iters), lml=float(fit.stats[1, -1]),
                pehe_train=float(np.sqrt(np.mean((tau[train] - t_tr["map"]) ** 2))),
                pehe_test=float(np.sqrt(np.mean((tau[test] - t_te["map"]) ** 2))),
                rmse_train=float(np.sqrt(np.mean((d["y"][train] - pred["map"][train]) ** 2))),
                rmse_test=float(np.sqrt(np.mean((d["y"][test] - pred["map"][test]) ** 2))),
                e_ate_train=float(np.mean(tau[train]) - t_tr["ate"]), e_ate_test=float(np.mean(tau[test]) - t_te["ate"]))


def hill_units_concurrent(units, n=747, p=25, test_frac=0.1, maxiter=3000, tol=1e-1, lr=0.01, prior=False, nu=2000,
                          batch=32, tile=0, covariates=None):
    """The same fits as `hill_unit`, but all `units` = [(replicate, restart), ...] are optimised
    together on this GPU, which is how the small n ~ 700 fits fill a B200: groups of `batch` fits share
    one batched handle (cs_fit_create_batch: every launch serves the whole group) and the groups are
    driven concurrently (cs_fit_run_many).  batch <= 1: one handle and stream set per fit.
    covariates=(X, z): draw the response surface on given covariates / treatment indicator (the real IHDP
    data of the reference's workspace files, causalstump_b200.rdata) instead of synthetic ones.
    Returns one result dict per unit."""
    import math as _m
    import time as _t
    from . import _lib, hill, main as M, rcpp_exports as rx
    lib = rx.lib()
    prep, fits, members = [], [], []
    t_start = _t.perf_counter()
    if covariates is not None:
        n, p = np.asarray(covariates[0]).shape
    for replicate, restart in units:
        if covariates is None:
            d = hill.hill_surface_b(n=n, p=p, replicate=replicate)
        else:
            d = hill.hill_surface_b(replicate=replicate, X=covariates[0], z=covariates[1])
        rng = np.random.default_rng(565 + 5 * replicate)
        ntest = int(_m.ceil(n * test_frac))
        test = np.sort(rng.choice(n, ntest, replace=False))
        train = np.setdiff1d(np.arange(n), test)
        nr = M.norm_variables(d["y"][train], d["X"][train])
        y, X, z = np.ascontiguousarray(nr["y"]), np.asfortranarray(nr["X"]), np.ascontiguousarray(d["z"][train])
        K = M.KernelClass_TP_SE(p, np.ones(len(y))) if prior else M.KernelClass_GP_SE(p, np.ones(len(y)))
        draws = hill.restart_draws(replicate, restart, prior)
        if prior:
            K.parainit(y, nu, draws)
        else:
            K.parainit(y, draws)
        if batch > 1:
            members.append((y, X, z, None, rx.pack(K.parameters)))
        else:
            fits.append(_lib.Fit(lib, K.kind, y, X, z, None, rx.pack(K.parameters), optim=_lib.CS_OPT_NADAM, lr=lr))
        kind = K.kind
        prep.append((d, nr["moments"], train, test, replicate, restart))
    t_prep = _t.perf_counter()
    if batch > 1:
        handles = [_lib.FitBatch(lib, kind, members[i:i + batch], optim=_lib.CS_OPT_NADAM, lr=lr, tile=tile)
                   for i in range(0, len(members), batch)]
        t_create = _t.perf_counter()
        res = _lib.run_many(handles, maxiter, tol)
        views = [(h, b) for h in handles for b in range(h.B)]
    else:
        handles = fits
        t_create = _t.perf_counter()
        res = _lib.run_many(fits, maxiter, tol)
        views = [(f, None) for f in fits]
    t_run = _t.perf_counter()
    out = []
    for (d, mom, train, test, replicate, restart), (f, member), (iters, trace) in zip(prep, views, res):
        if member is not None:
            f.select(member)
        sd = _m.sqrt(mom["varY"])
        Xall = M.norm_variables(X=d["X"], moments=mom)["X"]
        pred = f.predict(Xall, d["z"])["map"] * sd + mom["meanY"]
        t_tr = f.treat(Xall[train])
        t_te = f.treat(Xall[test])
        tau = d["tau"]
        out.append(dict(replicate=replicate, restart=restart, iters=int(iters), lml=float(trace[1, -1]),
                        pehe_train=float(np.sqrt(np.mean((tau[train] - sd * t_tr["map"]) ** 2))),
                        pehe_test=float(np.sqrt(np.mean((tau[test] - sd * t_te["map"]) ** 2))),
                        rmse_train=float(np.sqrt(np.mean((d["y"][train] - pred[train]) ** 2))),
                        rmse_test=float(np.sqrt(np.mean((d["y"][test] - pred[test]) ** 2))),
                        e_ate_train=float(np.mean(tau[train]) - sd * t_tr["ate"]),
                        e_ate_test=float(np.mean(tau[test]) - sd * t_te["ate"])))
    for h in handles:
        h.close()
    LAST_TIMING.update(prep=t_prep - t_start, create=t_create - t_prep, run=t_run - t_create, post=_t.perf_counter() - t_run)
    return out


LAST_TIMING = {}  # seconds spent by the last hill_units_concurrent call: host data prep / handle creation / device loop / predict+treat


def select_winners(results):
    """Best restart (highest final LML) per replicate, ordered by replicate."""
    best = {}
    for r in results:
        k = r["replicate"]
        if k not in best or r["lml"] > best[k]["lml"] or (r["lml"] == best[k]["lml"] and r["restart"] < best[k]["restart"]):
            best[k] = r
    return [best[k] for k in sorted(best)]


def run_sharded(unit_fn, n_replicates, n_restarts, rank=0, world=1, dist=None):
    """Runs this rank's share of the units; returns (all results on rank 0 | None, local results)."""
    units = unit_list(n_replicates, n_restarts)
    local = [unit_fn(rep, rs) for rep, rs in shard(units, rank, world)]
    if world == 1 or dist is None:
        return sorted(local, key=lambda r: (r["replicate"], r["restart"])), local
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank != 0:
        return None, local
    allr = [r for part in gathered for r in part]
    return sorted(allr, key=lambda r: (r["replicate"], r["restart"])), local
