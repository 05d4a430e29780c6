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


METRICS = ("PEHE.all.train", "PEHE.tt.train", "RMSE.all.train", "RMSE.tt.train", "E.ate.train", "E.att.train", "SE.ate.train",
           "SE.att.train", "coverage.ite.train", "coverage.ate.train", "range.ate.train", "PEHE.all.test", "PEHE.tt.test",
           "RMSE.all.test", "RMSE.tt.test", "E.ate.test", "E.att.test", "SE.ate.test", "SE.att.test", "coverage.ite.test",
           "coverage.ate.test", "range.ate.test")


def update_results(testindex, y_hat_all, tau_hat_all, tau_ci_all, ate_ci_train, ate_ci_test, y, z, y1, y0):
    """The metric block of the simulation scripts, /root/reference/test/sec5-3_Hill2011.R:104-152, as a dict over
    METRICS (row order of `result.item`, :163-164).  Reproduced as written, including the `[z.train=1]` subscripts
    of the PEHE.tt rows (:122,:124): `x[z.train=1]` passes a named argument, not a comparison, so R selects the
    FIRST element -- PEHE.tt.* is |error of unit 1|, while E.att / RMSE.tt use the intended `==` masks."""
    n = len(y)
    te = np.zeros(n, dtype=bool)
    te[np.asarray(testindex)] = True
    tr = ~te
    tau = np.asarray(y1) - np.asarray(y0)
    out = {}
    for name, m in (("train", tr), ("test", te)):
        err = (tau - tau_hat_all)[m]
        zz, yy, yh = np.asarray(z)[m], np.asarray(y)[m], np.asarray(y_hat_all)[m]
        t1 = zz == 1
        out["PEHE.all." + name] = float(np.sqrt(np.mean(err ** 2)))
        out["PEHE.tt." + name] = float(np.sqrt(np.mean(err[:1] ** 2)))  # quirk, see above
        e_ate = float(np.mean(tau[m]) - np.mean(tau_hat_all[m]))
        e_att = float(np.mean(tau[m][t1]) - np.mean(tau_hat_all[m][t1])) if t1.any() else float("nan")
        out["E.ate." + name], out["E.att." + name] = e_ate, e_att
        out["SE.ate." + name], out["SE.att." + name] = e_ate ** 2, e_att ** 2
        out["RMSE.all." + name] = float(np.sqrt(np.mean((yy - yh) ** 2)))
        out["RMSE.tt." + name] = float(np.sqrt(np.mean(((yy - yh)[t1]) ** 2))) if t1.any() else float("nan")
        ci = np.asarray(tau_ci_all)[m]
        out["coverage.ite." + name] = float(np.mean((ci[:, 0] < tau[m]) & (ci[:, 1] > tau[m])))
        ac = ate_ci_train if name == "train" else ate_ci_test
        out["coverage.ate." + name] = float((ac[0] < np.mean(tau[m])) and (ac[1] > np.mean(tau[m])))
        out["range.ate." + name] = float(ac[1] - ac[0])
    return out


def hill_units_concurrent(units, n=747, p=25, test_frac=0.1, maxiter=3000, tol=1e-1, lr=0.01, prior=False, nu=2000,
                          batch=32, tile=0, covariates=None, chunk=0):
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
        handles = [_lib.FitBatch(lib, kind, members[i:i + batch], optim=_lib.CS_OPT_NADAM, lr=lr, tile=tile, chunk=chunk)
                   for i in range(0, len(members), batch)]
        t_create = _t.perf_counter()
        res = _lib.run_many(handles, maxiter, tol, strict=False)
        views = [(h, b) for h in handles for b in range(h.B)]
    else:
        handles = fits
        t_create = _t.perf_counter()
        res = _lib.run_many(fits, maxiter, tol, strict=False)
        views = [(f, None) for f in fits]
    # a unit whose K_n stopped being positive definite (or whose LML became NaN) fails ALONE, as the one CausalStump() call
    # of that replicate would in the reference; its row says so and the other units of the shard are unaffected
    status = [st for h in handles for st in h.member_status()]
    t_run = _t.perf_counter()
    out = []
    for (d, mom, train, test, replicate, restart), (f, member), (iters, trace), stt in zip(prep, views, res, status):
        if stt != 0:
            nan = float("nan")
            out.append(dict(replicate=replicate, restart=restart, iters=int(iters), lml=-float("inf"), failed=True, status=int(stt),
                            metrics={k: nan for k in METRICS}, pehe_train=nan, pehe_test=nan, rmse_train=nan, rmse_test=nan,
                            e_ate_train=nan, e_ate_test=nan))
            continue
        if member is not None:
            f.select(member)
        sd = _m.sqrt(mom["varY"])
        Xall = M.norm_variables(X=d["X"], moments=mom)["X"]
        pred = f.predict(Xall, d["z"])["map"] * sd + mom["meanY"]
        t_tr = f.treat(Xall[train])
        t_te = f.treat(Xall[test])
        tau = d["tau"]
        # the 22 metrics of update.results; intervals as treatment.R:28-31 builds them for the GP class
        # (varY * (+-1.96 * variance) + map, quirk Q8; ATE interval from sqrt(ate_var))
        vy = mom["varY"]
        tau_hat, tau_ci = np.empty(n), np.empty((n, 2))
        for idx, t in ((train, t_tr), (test, t_te)):
            tau_hat[idx] = sd * t["map"]
            tau_ci[idx, 0] = vy * (-1.96 * t["var"]) + sd * t["map"]
            tau_ci[idx, 1] = vy * (1.96 * t["var"]) + sd * t["map"]
        # sqrt of a negative ate_var is NaN as in R (R/treatment.R:31): a broken fit must not show up as a zero-width interval
        with np.errstate(invalid="ignore"):
            half = [float(np.sqrt(np.float64(t["ate_var"]))) for t in (t_tr, t_te)]
        ate_ci = [(vy * (-1.96 * h) + sd * t["ate"], vy * (1.96 * h) + sd * t["ate"]) for h, t in zip(half, (t_tr, t_te))]
        metrics = update_results(test, pred, tau_hat, tau_ci, ate_ci[0], ate_ci[1], d["y"], d["z"], d["y1"], d["y0"])
        out.append(dict(replicate=replicate, restart=restart, iters=int(iters), lml=float(trace[1, -1]), metrics=metrics, failed=False, status=0,
                        pehe_train=float(np.sqrt(np.mean((tau[train] - sd * t_tr["map"]) ** 2))),
                        pehe_test=float(np.sqrt(np.mean((tau[test] - sd * t_te["map"]) ** 2))),
                        rmse_train=float(np.sqrt(np.mean((d["y"][train] - pred[train]) ** 2))),
                        rmse_test=float(np.sqrt(np.mean((d["y"][test] - pred[test]) ** 2))),
                        e_ate_train=float(np.mean(tau[train]) - sd * t_tr["ate"]),
                        e_ate_test=float(np.mean(tau[test]) - sd * t_te["ate"])))
    for h in handles:
        h.close()
    # slot utilisation of the batched handles: a member that has converged keeps its blockIdx.z slot (as early-exit CTAs)
    # until the slowest member of ITS handle stops -- sum(iterations) / sum_h(B_h * max iterations in h)
    used = cap = 0
    k = 0
    for h in handles:
        its = [res[k + b][0] for b in range(getattr(h, "B", 1))]
        k += len(its)
        used += sum(its); cap += len(its) * max(its)
    LAST_TIMING.update(prep=t_prep - t_start, create=t_create - t_prep, run=t_run - t_create, post=_t.perf_counter() - t_run,
                       slot_utilisation=used / max(cap, 1), iterations=used, failed=sum(1 for s in status if s != 0))
    return out


def hill_units_device(units, n=747, p=25, test_frac=0.1, maxiter=3000, tol=1e-1, lr=0.01, batch=32, covariates=None, seed=565):
    """The replicate loop with NOTHING but the unit list going in and the metric block coming out (cs_sim_run, csrc/sim.cuh):
    response surface B, the 10 % split, norm_variables and parainit are generated on the device straight into the members of
    batched fit handles, the fits run as in `hill_units_concurrent`, predict / treatment keep their inputs and outputs on the
    device, and update.results (:104-152) is one kernel per batch.  Draws come from Philox counters (host mirror:
    hill.hill_unit_philox), so a unit is NOT the same data as hill.hill_surface_b(replicate) -- same law, different stream.
    Returns the same per-unit dicts as `hill_units_concurrent` (GP arm)."""
    import time as _t
    from . import rcpp_exports as rx
    lib = rx.lib()
    t0 = _t.perf_counter()
    X, z = (None, None) if covariates is None else covariates
    out = lib.sim_run(units, n=n, p=p, X=X, z=z, test_frac=test_frac, maxiter=maxiter, tol=tol, lr=lr, batch=batch, seed=seed)
    res = []
    for u, (replicate, restart) in enumerate(units):
        m = dict(zip(METRICS, out["metrics"][u].tolist()))
        res.append(dict(replicate=replicate, restart=restart, iters=int(out["iters"][u]), lml=float(out["lml"][u]), metrics=m,
                        failed=bool(out["status"][u] != 0), status=int(out["status"][u]), pehe_train=m["PEHE.all.train"],
                        pehe_test=m["PEHE.all.test"], rmse_train=m["RMSE.all.train"], rmse_test=m["RMSE.all.test"],
                        e_ate_train=m["E.ate.train"], e_ate_test=m["E.ate.test"]))
    used = cap = 0
    for i in range(0, len(units), batch):
        its = [int(x) for x in out["iters"][i:i + batch]]
        used += sum(its); cap += len(its) * max(its)
    LAST_TIMING.clear()
    LAST_TIMING.update(total=_t.perf_counter() - t0, slot_utilisation=used / max(cap, 1), iterations=used,
                       failed=int(np.sum(out["status"] != 0)), prep=0.0, post=0.0)
    return res


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
