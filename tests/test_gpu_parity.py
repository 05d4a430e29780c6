"""Parity tests proper: the CUDA library on a B200, through the C ABI, against the oracle
on the same seeded inputs (sizes the oracle finishes in seconds), against the committed
golden fixtures, and -- at the metric's full size n=8192 -- through size-independent
properties.  Tolerances are BASELINE.json's: LML 1e-9 rel, gradients / predictive means
1e-8 rel, PEHE 6 significant digits after the same number of optimizer steps."""
import os

import numpy as np
import pytest

from causalstump_b200 import _lib, hill
from oracle import causalstump_oracle as o
from tests import _cases as cs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n,p,tile,kw", [
    ("GP", 120, 1, 0, dict(readme=True)),            # config 1 (README)
    ("GP", 747, 25, 0, {}),                          # config 2 (Hill, n=747)
    ("GP", 672, 26, 0, dict(weights=True)),          # training split + propensity column, w=
    ("TP", 747, 25, 0, dict(nu=2000.0)),             # config 3
    ("TP", 300, 5, 128, dict(nu=2.0)),
    ("GP", 1000, 25, 128, {}),                       # 128-tile engine, ragged n
    ("GP", 64, 3, 64, {}), ("GP", 65, 3, 64, {}), ("GP", 2, 2, 0, {}), ("GP", 3, 1, 128, {}),  # edge sizes
])
def test_eval_parity(gpulib, kind, n, p, tile, kw):
    cs.check_eval(gpulib, kind, n, p, tile=tile, **kw)


@pytest.mark.parametrize("kind,n,p,optim,iters,kw", [
    ("GP", 120, 1, "Nadam", 25, dict(readme=True)),
    ("GP", 400, 25, "Nadam", 12, {}),
    ("TP", 400, 26, "Nadam", 12, dict(nu=2000.0)),
    ("GP", 300, 5, "Adam", 12, dict(weights=True)),
    ("GP", 200, 3, "Nesterov", 10, {}),
    ("GP", 200, 3, "GD", 10, {}),
])
def test_fit_predict_treat_pehe_parity(gpulib, kind, n, p, optim, iters, kw):
    cs.check_fit(gpulib, kind, n, p, optim=optim, maxiter=iters, n2=75, **kw)


def test_chunking_is_bitwise_invariant(gpulib):
    P = cs.make_problem("GP", 200, 4)
    outs = []
    for chunk in (1, 7, 64):
        f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), chunk=chunk)
        it, tr = f.run(15, 1e-12)
        outs.append((it, tr.copy(), f.get_params()))
        f.close()
    for it, tr, pv in outs[1:]:
        assert it == outs[0][0] and np.array_equal(tr, outs[0][1]) and np.array_equal(pv, outs[0][2])


def test_early_stop_matches_oracle(gpulib):
    P = cs.make_problem("GP", 150, 3)
    fit = cs.oracle_fit(P, "Nadam", 60, 2.0, 0.01)
    f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]))
    it, trace = f.run(60, 2.0)
    f.close()
    assert it == fit.iters and cs.rel(trace[1, 1:], fit.stats[1, 1:]) < 1e-9


def test_golden_fixtures(gpulib):
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
    for nm in sorted({k.split("/")[0] for k in G.files}):
        kind, n, p, iters, optim = (int(v) for v in G[nm + "/meta"])
        f = _lib.Fit(gpulib, kind, G[nm + "/y"], G[nm + "/X"], G[nm + "/z"], G[nm + "/w"], G[nm + "/init_params"],
                     optim=optim, lr=0.02, momentum=0.3)
        g, st, mu = f.eval()
        gref = G[nm + "/grad"]
        assert np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD
        assert abs(st[1] - G[nm + "/stats"][1]) < cs.TOL_LML * abs(G[nm + "/stats"][1])
        assert abs(mu - float(G[nm + "/mu_sol"])) < 1e-9 * abs(float(G[nm + "/mu_sol"]))
        it, trace = f.run(iters, 1e-14)
        assert it == iters
        assert cs.rel(f.get_params(), G[nm + "/fit_params"], 1e-9) < 1e-7
        assert cs.rel(trace[1, 1:], G[nm + "/fit_trace"][1, 1:]) < 1e-8
        jit = 1e-6 if kind == 1 else 0.0
        pr = f.predict(G[nm + "/X2"], G[nm + "/z2"])
        tr = f.treat(G[nm + "/X2"], cov_jitter=jit)
        assert cs.relm(pr["map"], G[nm + "/pred_map"]) < cs.TOL_MAP
        assert cs.relm(pr["var"], G[nm + "/pred_var"]) < 1e-7
        assert cs.relm(tr["map"], G[nm + "/treat_map"]) < cs.TOL_MAP
        assert abs(tr["ate"] - float(G[nm + "/treat_ate"])) < 1e-8 * max(abs(float(G[nm + "/treat_ate"])), 1e-3)
        assert abs(tr["ate_var"] - float(G[nm + "/treat_ate_var"])) < 1e-7 * abs(float(G[nm + "/treat_ate_var"]))
        f.close()


def test_golden_ref(gpulib):
    """CUDA path vs vectors produced by the reference's own C++ (oracle/_ref; tests/golden/make_golden_ref.py)."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref_v1.npz"))
    for nm in sorted({k.split("/")[0] for k in G.files}):
        kind, n, p, iters, optim = (int(v) for v in G[nm + "/meta"][:5])
        lr, mom = float(G[nm + "/meta"][5]), float(G[nm + "/meta"][6])
        f = _lib.Fit(gpulib, kind, G[nm + "/y"], G[nm + "/X"], G[nm + "/z"], G[nm + "/w"], G[nm + "/init_params"],
                     optim=optim, lr=lr, momentum=mom)
        g, st, mu = f.eval()
        gref = G[nm + "/grad"]
        assert np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD, nm
        assert abs(st[1] - G[nm + "/stats"][1]) < cs.TOL_LML * abs(G[nm + "/stats"][1]), nm
        assert abs(st[0] - G[nm + "/stats"][0]) < 1e-7 * abs(G[nm + "/stats"][0]), nm
        assert abs(mu - float(G[nm + "/mu_sol"])) < 1e-9 * abs(float(G[nm + "/mu_sol"])), nm
        m = f.get_matrices(want=("Kmat", "Km", "Ka", "invK"))
        assert cs.relm(m["Kmat"][0], G[nm + "/K_row0"]) < 1e-13 and cs.relm(m["Km"][0], G[nm + "/Km_row0"]) < 1e-13
        assert cs.relm(m["Ka"][0], G[nm + "/Ka_row0"]) < 1e-13
        assert cs.relm(np.diag(m["invK"]), G[nm + "/invK_diag"]) < 1e-9 and cs.relm(m["invK"][0], G[nm + "/invK_row0"]) < 1e-9
        it, trace = f.run(iters, 0.0)
        assert it == iters
        assert cs.rel(f.get_params(), G[nm + "/fit_params"], 1e-9) < 1e-7, nm
        assert cs.rel(trace[1, 1:], G[nm + "/fit_trace"][1, 1:]) < 1e-8, nm
        assert cs.rel(trace[0, 1:], G[nm + "/fit_trace"][0, 1:]) < 1e-6, nm
        pr = f.predict(G[nm + "/X2"], G[nm + "/z2"])
        tr = f.treat(G[nm + "/X2"], cov_jitter=1e-6 if kind == 1 else 0.0)
        assert cs.relm(pr["map"], G[nm + "/pred_map"]) < cs.TOL_MAP and cs.relm(pr["var"], G[nm + "/pred_var"]) < 1e-7, nm
        assert cs.relm(tr["map"], G[nm + "/treat_map"]) < cs.TOL_MAP and cs.relm(tr["var"], G[nm + "/treat_var"]) < 1e-7, nm
        assert abs(tr["ate"] - float(G[nm + "/treat_ate"])) < 1e-8 * max(abs(float(G[nm + "/treat_ate"])), 1e-3), nm
        assert abs(tr["ate_var"] - float(G[nm + "/treat_ate_var"])) < 1e-7 * abs(float(G[nm + "/treat_ate_var"])), nm
        f.close()


def test_limits_and_failure_reporting(gpulib):
    cs.check_limits(gpulib)


def test_failing_member_stops_alone_on_gpu(gpulib):
    """same body as the emulator test: a non positive definite member and a NaN member stop alone, the healthy member of the
    batch finishes bit-identically to a single handle (device-side WHILE loop included)"""
    cs.check_failing_member(gpulib)


def test_batched_handle(gpulib):
    """cs_fit_create_batch at the size of config 4 (672 x 25 training splits): all members advance with every
    launch; each must equal its own single-handle fit (to rounding: the per-fit CTA count of the trace-gradient
    reduction differs between a batch and a single handle) and the oracle to tolerance."""
    B, n, p = 6, 672, 25
    Ps = [cs.make_problem("GP", n, p, replicate=r, draws=(0.3 + 0.1 * r, 0.9 - 0.1 * r)) for r in range(B)]
    members = [(P["y"], P["X"], P["z"], None, o.pack_params(P["par"])) for P in Ps]
    fb = _lib.FitBatch(gpulib, 0, members)
    fb.eval_async(); fb.sync()
    for b in (0, B - 1):
        P = Ps[b]
        g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
        g2, st2, mu2 = fb.select(b).fetch()
        gref = o.pack_params(g)
        assert np.max(np.abs(g2 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD
        assert abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1]) and abs(mu2 - mu) < 1e-9 * abs(mu)
    res = fb.select(0).run(40, 2.0)
    its = [it for it, _ in res]
    for b, P in enumerate(Ps):
        f1 = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], None, o.pack_params(P["par"]))
        it1, tr1 = f1.run(40, 2.0)
        assert its[b] == it1 and cs.rel(res[b][1][1, 1:], tr1[1, 1:]) < 1e-11
        assert cs.rel(fb.select(b).get_params(), f1.get_params(), 1e-9) < 1e-9
        X2 = P["X"][:75] * 0.97
        assert cs.relm(fb.treat(X2)["map"], f1.treat(X2)["map"]) < 1e-9
        assert cs.relm(fb.predict(X2, P["z"][:75])["map"], f1.predict(X2, P["z"][:75])["map"]) < 1e-9
        f1.close()
    ref = cs.oracle_fit(Ps[1], "Nadam", 40, 2.0, 0.01)
    assert its[1] == ref.iters and cs.rel(res[1][1][1, 1:], ref.stats[1, 1:]) < 1e-8
    fb.close()


def test_real_ihdp_covariates(gpulib):
    """The reference's own data: IHDP covariates and treatment indicator decoded from its workspace file by
    causalstump_b200/rdata.py (tests/golden/ihdp_covariates.npz), response surface B drawn on top.  Evaluation
    parity against the oracle at n = 747, p = 25, then a 12-replicate mini simulation through the batched driver
    whose aggregates must sit inside the band of the GP arm stored in the reference's `results` matrix."""
    from causalstump_b200 import replicates as rp
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ihdp_covariates.npz"))
    X, z = G["X"], G["Trt"]
    d = hill.hill_surface_b(replicate=3, X=X, z=z)
    nr = o.norm_variables(d["y"], d["X"])
    K = o.KernelClass_GP_SE(25, np.ones(747))
    K.parainit(nr["y"], 0.5, 0.7)
    g, st, mu = o.eval_lml_grad("GP", nr["y"], nr["X"], d["z"], np.ones(747), K.parameters)
    f = _lib.Fit(gpulib, 0, nr["y"], np.asfortranarray(nr["X"]), d["z"], None, o.pack_params(K.parameters))
    g2, st2, mu2 = f.eval()
    f.close()
    gref = o.pack_params(g)
    assert np.max(np.abs(g2 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD
    assert abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1]) and abs(mu2 - mu) < 1e-9 * abs(mu)
    res = rp.hill_units_concurrent([(r, 0) for r in range(12)], covariates=(X, z), batch=12)
    ref = dict(zip(G["metrics"], G["gp_results"]))
    pehe = np.mean([r["pehe_train"] for r in res])
    rmse = np.mean([r["rmse_train"] for r in res])
    lo, hi = np.nanpercentile(ref["PEHE.all.train"], [2, 98])
    assert 0.5 * lo < pehe < max(hi, 3.0), (pehe, lo, hi)          # reference: mean 1.40, heavy right tail
    assert 0.6 < rmse < 1.3, rmse                                   # reference: 0.88 (noise sd 1)
    assert all(3 < r["iters"] <= 3000 for r in res)
    # the full metric block (update.results) against the stored GP arm: same quirky interval construction
    M = {k: np.mean([r["metrics"][k] for r in res]) for k in rp.METRICS}
    R = {k: np.nanmean(v) for k, v in ref.items()}
    assert abs(M["PEHE.all.train"] - pehe) < 1e-12 and M["coverage.ate.train"] == 1.0 == R["coverage.ate.train"]
    assert 0.4 < M["coverage.ite.train"] < 0.95 and abs(R["coverage.ite.train"] - 0.67) < 0.01
    assert 0.5 * R["range.ate.train"] < M["range.ate.train"] < 2.0 * R["range.ate.train"], (M["range.ate.train"], R["range.ate.train"])
    assert 0.4 * R["range.ate.test"] < M["range.ate.test"] < 2.5 * R["range.ate.test"], (M["range.ate.test"], R["range.ate.test"])
    assert abs(M["E.ate.train"]) < 0.6 and 0.7 < M["RMSE.all.test"] < 2.5


def test_r_glue_on_the_cuda_library(gpulib, tmp_path):
    """tests/test_r_glue.py with the glue linked to the CUDA library instead of the emulated one: the reference's 11 .Call
    routines and the b200_* fast path, re-implemented on the C ABI, against the reference's own compiled routines."""
    import shutil
    from oracle import ref_binding as rb
    from tests import test_r_glue as trg
    if not shutil.which("g++") or not rb.available():
        pytest.skip("g++ or oracle/_ref/libcausalstump_ref.so not present")
    out = rb.build_glue(_lib.DEFAULT_LIB, str(tmp_path / "libglue_cuda.so"))
    libs = (rb.RefLib(), rb.RefLib(out, prefix="glue"))
    trg.test_glue_routines_equal_the_reference_routines(libs, "GP", 300, 25, True)
    trg.test_glue_routines_equal_the_reference_routines(libs, "TP", 130, 4, False)
    for optim in (0, 1, 2):
        trg.test_glue_optimizers_equal_the_reference_optimizers(libs, optim)
    trg.test_glue_fast_path_equals_the_restated_r_loop(libs)


def test_stateless_mirrors_vs_compiled_reference(gpulib):
    """The stateless C-ABI mirrors against oracle/_ref (the reference's C++ compiled with the header shim),
    which travels to the GPU box as a prebuilt .so; skipped when it is absent."""
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libcausalstump_ref.so not present")
    ref = rb.RefLib()
    for kind, n, p in ((0, 300, 25), (1, 257, 7)):
        P = cs.make_problem("GP" if kind == 0 else "TP", n, p, weights=True)
        par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
        flat = o.pack_params(par)
        la = par["lambdaa"] if kind == 0 else par["lambda0"]
        K, Km, Ka = ref.kernmat(kind, X[:101], X, z[:101], z, par["Lm"], par["La"], par["lambdam"], la)
        out = gpulib.kernmat(X[:101], X, z[:101], z, par["Lm"], par["La"], par["lambdam"], la)
        assert cs.relm(out["K"], K) < 1e-13 and cs.relm(out["Km"], Km) < 1e-13 and cs.relm(out["Ka"], Ka) < 1e-13
        g, st, mu, mats = ref.eval_lml_grad(kind, y, X, z, w, flat)
        inv2 = gpulib.invkernel(w, mats["K"], par["sigma"])
        assert cs.relm(inv2, mats["invK"]) < 1e-9
        assert abs(gpulib.mu_solution(y, mats["invK"]) - mu) < 1e-10 * abs(mu)
        for kw in ({}, dict(invK=mats["invK"], Kmat=mats["K"], Km=mats["Km"], Ka=mats["Ka"])):
            g2, st2 = gpulib.grad(kind, y, X, z, w, flat, **kw)
            assert np.max(np.abs(g2 - g) / np.maximum(np.abs(g), 1e-6 * np.max(np.abs(g)))) < cs.TOL_GRAD
            assert abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1]) and abs(st2[0] - st[0]) < 1e-7 * abs(st[0])
        m0, v0 = np.zeros_like(flat), np.zeros_like(flat)
        for optim in (0, 1, 2):
            mr, vr, pr_ = ref.opt_step(optim, kind, p, 3.0, 0.01, 0.9 if optim < 2 else 0.5, 0.999, 1e-8, m0 + 0.01, v0 + 0.02, g, flat)
            m2, v2, p2 = gpulib.opt_step(optim, 3.0, 0.01, 0.9 if optim < 2 else 0.5, 0.999, 1e-8, m0 + 0.01, v0 + 0.02, g, flat)
            # the C-ABI step leaves mu to the caller's closed form (quirk Q1): compare all entries but the last
            assert np.allclose(p2[:-1], pr_[:-1], rtol=1e-13, atol=0) and np.allclose(m2[:-1], mr[:-1], rtol=1e-13, atol=1e-300)


def test_stateless_mirrors(gpulib):
    P = cs.make_problem("GP", 500, 25, weights=True)
    par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
    Kl = list(o.kernmat_GP_SE_cpp(X[:123], X, z[:123], z, par).values())
    out = gpulib.kernmat(X[:123], X, z[:123], z, par["Lm"], par["La"], par["lambdam"], par["lambdaa"])
    assert cs.relm(out["K"], Kl[0]) < 1e-13 and cs.relm(out["Km"], Kl[1]) < 1e-13 and cs.relm(out["Ka"], Kl[2]) < 1e-13
    Kt = list(o.kernmat_GP_SE_cpp(X, X, z, z, par).values())
    inv = o.invkernel_cpp(z, w, Kt[0], par)
    inv2, ld = gpulib.invkernel(w, Kt[0], par["sigma"], want_logdet=True)
    assert cs.relm(inv2, inv) < 1e-10 and abs(ld + o.log_det(inv)[0]) < 1e-10 * abs(ld)
    assert abs(gpulib.mu_solution(y, inv) - o.mu_solution_cpp(y, z, inv, par)) < 1e-11
    gl = o.grad_GP_SE_cpp(y, X, z, w, Kt[0], Kt[1], Kt[2], inv, par)
    gref = o.pack_params(gl["gradients"])
    for kw in ({}, dict(invK=inv), dict(invK=inv, Kmat=Kt[0], Km=Kt[1], Ka=Kt[2])):
        g, st = gpulib.grad(0, y, X, z, w, o.pack_params(par), **kw)
        assert np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD
        assert abs(st[1] - gl["stats"][1]) < cs.TOL_LML * abs(gl["stats"][1])
    st = gpulib.stats(0, y, Kt[0], inv, par["mu"])
    assert cs.rel(st, o.stats_GP_SE(y, Kt[0], inv, par)) < 1e-9
    with pytest.raises(_lib.NotPositiveDefinite):
        gpulib.invkernel(None, -np.eye(70), 0.0)


def test_optimizer_mirrors(gpulib):
    # element-wise updates agree to rounding (the device contracts a*b+c into FMAs)
    from collections import OrderedDict
    rng = np.random.default_rng(1)
    m, v, g, pa = rng.normal(size=56), rng.uniform(size=56), rng.normal(size=56), rng.normal(size=56)

    def od(a):
        return OrderedDict(s=float(a[0]), lm=float(a[1]), la=float(a[2]), Lm=a[3:29].copy(), La=a[29:55].copy(), mu=float(a[55]))

    for code, fn in ((0, o.Adam_cpp), (1, o.Nadam_cpp)):
        r = fn(17, 0.01, 0.9, 0.999, 1e-8, od(m), od(v), od(g), od(pa))
        m2, v2, p2 = gpulib.opt_step(code, 17, 0.01, 0.9, 0.999, 1e-8, m, v, g, pa)
        assert cs.rel(m2, o.pack_params(r["m"])) < 1e-13 and cs.rel(v2, o.pack_params(r["v"])) < 1e-13
        assert cs.rel(p2, o.pack_params(r["parameters"])) < 1e-13
    r = o.Nesterov_cpp(0.01, 0.5, od(m), od(g), od(pa))
    m2, _, p2 = gpulib.opt_step(2, 1, 0.01, 0.5, 0, 0, m, None, g, pa)
    assert cs.rel(m2, o.pack_params(r["nu"])) < 1e-13 and np.array_equal(p2, o.pack_params(r["parameters"]))


def test_public_api_end_to_end(gpulib):
    """The call a user of the reference makes: CausalStump / predict / treatment."""
    import causalstump_b200 as cb
    d = hill.readme_example(120)
    for prior in (False, True):
        ofit = o.CausalStump(d["y"], d["X"], d["z"], maxiter=30, tol=1e-12, prior=prior, nu=200,
                             init_draws=(0.5, 0.4), faithful=False)
        cfit = cb.CausalStump(d["y"], d["X"], d["z"], maxiter=30, tol=1e-12, prior=prior, nu=200,
                              init_draws=(0.5, 0.4), verbose=False)
        assert cfit.iters == ofit.iters
        po, pc = o.predict(ofit, z=1), cb.predict(cfit, z=1)
        assert cs.relm(pc["map"], po["map"]) < cs.TOL_MAP
        to, tc = o.treatment(ofit), cb.treatment(cfit)
        assert cs.relm(tc["map"], to["map"]) < cs.TOL_MAP and abs(tc["ate"] - to["ate"]) < 1e-8 * abs(to["ate"])
        pehe_o = np.sqrt(np.mean((d["tau"] - to["map"]) ** 2))
        pehe_c = np.sqrt(np.mean((d["tau"] - tc["map"]) ** 2))
        assert abs(pehe_c - pehe_o) < 0.5e-6 * pehe_o
        if not prior:
            assert cs.relm(pc["ci"], o.predict(ofit, z=1)["ci"]) < 1e-7
            assert cs.relm(tc["ate_ci"], to["ate_ci"]) < 1e-7
            assert cs.relm(cfit.Kernel.invKmatn, ofit.Kernel.invKmatn) < 1e-9
        else:
            # sampling intervals: distributional parity only (mvnfast's RNG stream is not reproducible)
            from scipy import stats as ss
            cov = ofit.Kernel.predict(ofit.train_data["y"], ofit.train_data["X"], ofit.train_data["z"],
                                      ofit.train_data["X"], np.ones(120))["cov"]
            U = (pc["ci"][:, 1] - pc["map"]) / ofit.moments["varY"]
            ref = np.sqrt(np.diag(cov)) * ss.t.ppf(0.9, 200.0)
            assert np.all(np.abs(U / ref - 1.0) < 0.08)


def _expected_order_stat(df, k=900, n=1000):
    """E[k-th of n order statistic] of a standard t_df variable (exact, by quadrature over the Beta(k, n-k+1) law
    of the uniform order statistic): what `mean_j sort(rmvt(1000, ...))[900]` estimates (R/kernel_TP_SE_class.R:90)."""
    from scipy import integrate, stats as ss
    b = ss.beta(k, n - k + 1)
    return integrate.quad(lambda u: ss.t.ppf(u, df) * b.pdf(u), b.ppf(1e-14), b.ppf(1 - 1e-14), limit=400,
                          epsabs=1e-13, epsrel=1e-12)[0]


@pytest.mark.parametrize("which,df,nsampling,vs_oracle", [("treat672", 2000.0, 400000, True), ("predict747", 2000.0, 100000, False),
                                                           ("treat672", 3.0, 200000, False)])
def test_tp_sampler_at_config3_size(gpulib, which, df, nsampling, vs_oracle):
    """cs_tp_sample (Philox + Box-Muller + Marsaglia-Tsang + bitonic order statistic on the device) at the sizes of
    config 3: the 672 x 672 treatment covariance of a Student-t fit (R/kernel_TP_SE_class.R:112-124, incl. ate_U) and the
    747 x 747 predictive covariance (:84-93).  mvnfast's sitmo stream is not reproducible, so parity is distributional:
    every U_i within 2 % of sqrt(cov_ii) * E[900th of 1000 t_df order statistic] (exact quadrature; Monte-Carlo error of
    the estimate is 0.2-0.4 % at these draw counts), ate_U within 2 % of the same law for the row means (a univariate
    t_df with scale sqrt(sum cov)/n2), and -- 400 000 draws on both sides -- within 2 % of the oracle's restatement of
    rmvt + apply(..., sort) run with NumPy's generator."""
    import bench
    tinp = bench.make_inputs(672, 25, 1, kind="TP")
    f = _lib.Fit(gpulib, _lib.CS_TP, tinp["y"], tinp["X"], tinp["z"], tinp["w"], tinp["params"])
    f.run(5, 0.0)
    if which == "treat672":
        cov = f.treat(tinp["X"], cov_jitter=1e-6, want_cov=True)["cov"]
    else:
        pinp = bench.make_inputs(747, 25, 2, kind="TP")
        cov = f.predict(pinp["X"], pinp["z"], want_cov=True)["cov"]
    f.close()
    n2 = cov.shape[0]
    assert n2 == int(which[-3:]) and np.allclose(cov, cov.T, rtol=0, atol=1e-12)
    cov = 0.5 * (cov + cov.T)
    U, aU = gpulib.tp_sample(cov, df, nsampling, 2026, want_ate=True)
    U2, aU2 = gpulib.tp_sample(cov, df, nsampling, 2026, want_ate=True)
    assert np.array_equal(U, U2) and aU == aU2                       # counter-based RNG: reproducible per seed
    e = _expected_order_stat(df)
    assert np.max(np.abs(U / (np.sqrt(np.diag(cov)) * e) - 1.0)) < 0.02
    assert abs(aU / (np.sqrt(cov.sum()) / n2 * e) - 1.0) < 0.02
    assert not np.array_equal(U, gpulib.tp_sample(cov, df, nsampling, 2027))
    if vs_oracle:
        K = o.KernelClass_TP_SE(25, np.ones(672), nsampling=nsampling)
        K.parameters = dict(nu=np.log(df))
        Uo, aUo = K._sample_U(cov, np.random.default_rng(11), want_ate=True)
        assert np.max(np.abs(U / Uo - 1.0)) < 0.02 and abs(aU / aUo - 1.0) < 0.02


def test_full_size_properties_n8192(gpulib):
    """n=8192, p=25 (the metric's size): the oracle is too slow here, so check
    size-independent properties: gradient == central finite difference of the LML,
    K^-1 consistency (K_n alpha = ybar), determinism, LML finite."""
    P = cs.make_problem("GP", 8192, 25)
    f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]))
    p0 = o.pack_params(P["par"])
    g, st, mu = f.eval(p0)
    g_b, st_b, _ = f.eval(p0)
    assert np.array_equal(g, g_b) and np.array_equal(st, st_b)  # deterministic reductions
    assert np.isfinite(st).all() and np.isfinite(g).all()
    lay_mu = len(p0) - 1
    for idx in (0, 1, 2, 3, 3 + 25, lay_mu):
        h = 1e-5
        pa, pb = p0.copy(), p0.copy()
        pa[idx] += h; pb[idx] -= h
        fd = (f.eval(pa)[1][1] - f.eval(pb)[1][1]) / (2 * h)
        assert abs(fd - g[idx]) <= 2e-5 * max(1.0, abs(g[idx])), (idx, fd, g[idx])
    # residual identity: stats[0] = || exp(sigma)/w * alpha ||^2 when K_n^-1 is exact
    f.eval(p0)
    pr = f.predict(P["X"][:256], P["z"][:256])
    # in-sample prediction = y - exp(sigma) * alpha  =>  finite and close to y in the bulk
    assert np.isfinite(pr["map"]).all() and np.isfinite(pr["var"]).all() and (pr["var"] > 0).all()
    f.close()


@pytest.mark.parametrize("n", [8192, 4000])
def test_bench_plan_at_full_size(gpulib, n):
    """The plan `bench.py` times (inverse pipelined behind the Cholesky on seven streams, chain on its own SM
    partition) against the sequential plan on plain streams, at the metric's size and at a ragged size that
    still enables the partition (padded n = 4096): same gradient / LML / mu / K^-1 to rounding, re-entrant,
    and K_n K^-1 = I on sampled columns."""
    P = cs.make_problem("GP", n, 25, weights=True)
    p0 = o.pack_params(P["par"])
    outs = []
    for pipe, part in ((1, 0), (0, -1)):
        f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], p0, inv_pipeline=pipe, partition=part)
        g, st, mu = f.eval()
        g2, st2, _ = f.eval()
        assert np.array_equal(g, g2) and np.array_equal(st, st2)
        outs.append((g, st, mu, f.get_matrices(want=("invK",))["invK"] if pipe else None, f))
    (g1, s1, m1, inv1, f1), (g0, s0, m0, _, f0) = outs
    assert np.max(np.abs(g1 - g0) / np.maximum(np.abs(g0), 1e-6 * np.max(np.abs(g0)))) < 1e-9
    assert abs(s1[1] - s0[1]) < 1e-11 * abs(s0[1]) and abs(m1 - m0) < 1e-10 * abs(m0) and abs(s1[0] - s0[0]) < 1e-8 * abs(s0[0])
    assert np.array_equal(inv1, inv1.T)
    cols = np.array([0, 1, 63, 64, 127, 128, 511, 512, n // 2, n - 129, n - 2, n - 1])
    par = P["par"]
    Kc = o.kernmat_GP_SE_cpp(P["X"], P["X"][cols], P["z"], P["z"][cols], par)["Kmat"]          # n x 12 columns of K
    Kc[cols, np.arange(len(cols))] += np.exp(par["sigma"]) / P["w"][cols]
    E = inv1 @ Kc                                                                              # = I[:, cols]
    E[cols, np.arange(len(cols))] -= 1.0
    assert np.max(np.abs(E)) < 1e-8, np.max(np.abs(E))
    f1.close(); f0.close()


def test_optimisation_loop_with_sm_partition_and_graph(gpulib):
    """cs_fit_run at a size that enables the SM partition: the captured CUDA graph spans the green-context streams;
    same trajectory as eager launches on plain priority streams."""
    P = cs.make_problem("GP", 4200, 10)
    p0 = o.pack_params(P["par"])
    outs = []
    for kw in (dict(partition=0, use_graph=0, inv_pipeline=1), dict(partition=-1, use_graph=-1, inv_pipeline=0)):
        f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], p0, **kw)
        it, tr = f.run(4, 0.0)
        outs.append((it, tr.copy(), f.get_params()))
        f.close()
    (i1, t1, q1), (i0, t0, q0) = outs
    assert i1 == i0 == 4 and cs.rel(t1[1, 1:], t0[1, 1:]) < 1e-10 and cs.rel(q1, q0, 1e-9) < 1e-9


def test_distributed_driver_simulated_on_one_gpu(gpulib):
    """Config 5 driver with all ranks of a 2 x 2 grid on this GPU (collectives = device copies);
    the NCCL run over real ranks is tools/dist_check.py under torchrun."""
    P = cs.make_problem("GP", 1000, 25, weights=True)
    g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
    for grid in ((2, 2), (2, 1)):
        g2, st2, mu2 = gpulib.dist_sim_eval(0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), grid[0], grid[1], 128)
        gref = o.pack_params(g)
        assert np.max(np.abs(g2 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD
        assert abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1])
        assert abs(mu2 - mu) < 1e-8 * abs(mu)


def test_concurrent_fits_on_gpu(gpulib):
    """cs_fit_run_many: several fits interleaved on separate streams give the same results as
    running them one after another (bitwise: the kernels are deterministic)."""
    Ps = [cs.make_problem("GP", 300, 25, replicate=r) for r in range(4)]
    seq = []
    for P in Ps:
        f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]))
        it, tr = f.run(20, 1e-12)
        seq.append((it, tr.copy(), f.get_params()))
        f.close()
    fits = [_lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"])) for P in Ps]
    res = _lib.run_many(fits, 20, 1e-12)
    for (it0, tr0, p0), f, (it, tr) in zip(seq, fits, res):
        assert it == it0 and np.array_equal(tr, tr0) and np.array_equal(f.get_params(), p0)
        f.close()


def test_plan_race_detector_on_gpu_build(gpulib):
    assert gpulib.dll.cs_debug_check_plan(8192, 128, 64, 1) == 0


@pytest.mark.parametrize("n,p,tile", [(747, 25, 0), (2000, 25, 128), (1100, 5, 64)])
def test_pipelined_inverse_matches_recursive(gpulib, n, p, tile):
    """cfg.inv_pipeline: the inverse trails the Cholesky on its own stream; same results as the
    recursive inverse (to rounding) and the same parity against the oracle."""
    P = cs.make_problem("GP", n, p, weights=True)
    outs = []
    for mode in (0, 1):
        f = _lib.Fit(gpulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=tile, inv_pipeline=mode)
        g, st, mu = f.eval()
        g2, st2, _ = f.eval()
        assert np.array_equal(g, g2) and np.array_equal(st, st2)  # deterministic, re-entrant
        outs.append((g, st, mu, f.get_matrices(want=("invK",))["invK"]))
        f.close()
    (g0, s0, m0, i0), (g1, s1, m1, i1) = outs
    assert np.max(np.abs(g1 - g0) / np.maximum(np.abs(g0), 1e-6 * np.max(np.abs(g0)))) < 1e-9
    assert abs(s1[1] - s0[1]) < 1e-11 * abs(s0[1]) and abs(m1 - m0) < 1e-10 * abs(m0)
    assert cs.relm(i1, i0) < 1e-10
    if n <= 1200:
        go, sto, muo = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
        gref = o.pack_params(go)
        assert np.max(np.abs(g1 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))) < cs.TOL_GRAD
        assert abs(s1[1] - sto[1]) < cs.TOL_LML * abs(sto[1])
