"""Kernel logic on the SIMT emulator (tests/emu): the same .cu sources compiled with
-DCS_EMU, every CUDA thread a fiber.  Checks tiling, fragment layouts, launch plans and
the device-side loop against the oracle at small n.  Not a product path."""
import numpy as np
import pytest

from causalstump_b200 import _lib
from oracle import causalstump_oracle as o
from tests import _cases as cs


@pytest.mark.parametrize("kind,n,p,tile", [("GP", 70, 2, 64), ("GP", 150, 25, 64), ("TP", 129, 4, 64),
                                           ("GP", 140, 3, 128), ("GP", 64, 1, 64)])
def test_eval_matches_oracle(emulib, kind, n, p, tile):
    cs.check_eval(emulib, kind, n, p, tile=tile, weights=(n % 2 == 0))


@pytest.mark.parametrize("kind,optim", [("GP", "Nadam"), ("TP", "Adam"), ("GP", "Nesterov"), ("GP", "GD")])
def test_device_loop_matches_oracle(emulib, kind, optim):
    cs.check_fit(emulib, kind, 100, 3, optim=optim, maxiter=7, chunk=3, n2=30)


def test_readme_example_shape(emulib):
    cs.check_fit(emulib, "GP", 120, 1, readme=True, maxiter=5, n2=50)


def test_early_stop_and_trace(emulib):
    P = cs.make_problem("GP", 80, 2)
    fit = cs.oracle_fit(P, "Nadam", 40, 5.0, 0.01)
    f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), chunk=4)
    it, trace = f.run(40, 5.0)
    f.close()
    assert it == fit.iters and 3 < it < 40
    assert cs.rel(trace[1, 1:], fit.stats[1, 1:]) < 1e-9


def test_not_positive_definite_is_reported(emulib):
    with pytest.raises(_lib.NotPositiveDefinite) as ei:
        emulib.invkernel(None, np.array([[1.0, 2.0], [2.0, 1.0]]), -40.0)
    assert ei.value.code == 2  # LAPACK-style 1-based failing pivot


def test_stateless_mirrors(emulib):
    P = cs.make_problem("TP", 90, 3, weights=True, nu=30.0)
    par, y, X, z, w = P["par"], P["y"], P["X"], P["z"], P["w"]
    Sl = list(o.kernmat_TP_SE_cpp(X[:31], X, z[:31], z, par).values())
    out = emulib.kernmat(X[:31], X, z[:31], z, par["Lm"], par["La"], par["lambdam"], par["lambda0"])
    assert cs.relm(out["K"], Sl[0]) < 1e-14 and cs.relm(out["Ka"], Sl[2]) < 1e-14
    St = list(o.kernmat_TP_SE_cpp(X, X, z, z, par).values())
    inv = o.invkernel_cpp(z, w, St[0], par)
    assert cs.relm(emulib.invkernel(w, St[0], par["sigma"]), inv) < 1e-12
    assert abs(emulib.mu_solution(y, inv) - o.mu_solution_cpp(y, z, inv, par)) < 1e-13
    gl = o.grad_TP_SE_cpp(y, X, z, w, St[0], St[1], St[2], inv, par)
    for kw in ({}, dict(invK=inv), dict(invK=inv, Kmat=St[0], Km=St[1], Ka=St[2])):
        g, st = emulib.grad(1, y, X, z, w, o.pack_params(par), **kw)
        assert cs.rel(g, o.pack_params(gl["gradients"]), 1e-9) < 1e-9 and cs.rel(st, gl["stats"]) < 1e-10
    st = emulib.stats(1, y, St[0], inv, par["mu"], np.exp(par["nu"]))
    assert cs.rel(st, o.stats_TP_SE(y, St[0], inv, par)) < 1e-11


@pytest.mark.parametrize("n2", [24, 100])
def test_sampler_distribution(emulib, n2):
    from scipy import stats as ss
    rng = np.random.default_rng(5)
    A = rng.normal(size=(n2, n2))
    cov = A @ A.T / n2 + 0.5 * np.eye(n2)
    U, aU = emulib.tp_sample(cov, 6.0, 4000, 11, want_ate=True)
    ref = np.sqrt(np.diag(cov)) * ss.t.ppf(0.9, 6.0)
    assert np.all(np.abs(U / ref - 1.0) < 0.08)
    assert abs(aU / (np.sqrt(cov.sum()) / n2 * ss.t.ppf(0.9, 6.0)) - 1.0) < 0.08


@pytest.mark.parametrize("n,tile,gt", [(8192, 128, 64), (8192, 128, 128), (3000, 128, 64), (1000, 64, 64), (704, 64, 64), (129, 128, 64)])
def test_multistream_cholesky_plan_has_no_races(emulib, n, tile, gt):
    """Happens-before check of the three-stream launch plan (stream order + event edges):
    every pair of launches with overlapping footprints and a write must be ordered."""
    assert emulib.dll.cs_debug_check_plan(n, tile, gt, 1) == 0


@pytest.mark.parametrize("tail", ["3,2"])
def test_chain_stream_takes_a_share_of_the_inverse(emulib, monkeypatch, tail):
    """Large single fits (n >= 4096) give the first tile columns of the K^-1 accumulation to the chain stream once the
    Cholesky is done (CS_TAIL_A, default 3 from the middle panel on) and split the trailing update over two streams
    (look-ahead of depth three); forced here at a size the emulator handles.  Tile ownership is static, so the result
    is bitwise the one of the plain plan."""
    P = cs.make_problem("GP", 1100, 3)
    p0 = o.pack_params(P["par"])
    for t in ("2", tail):
        monkeypatch.setenv("CS_TAIL_A", t)
        assert emulib.dll.cs_debug_check_plan(3000, 128, 64, 1) == 0
    monkeypatch.setenv("CS_TAIL_A", tail)
    assert emulib.dll.cs_debug_check_plan(1536, 128, 64, 1) == 0 and emulib.dll.cs_debug_check_plan(2048, 128, 64, 1) == 0
    f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], p0, tile=128, inv_pipeline=1)
    g1, st1, mu1 = f.eval()
    f.close()
    monkeypatch.setenv("CS_TAIL_A", "0")
    monkeypatch.setenv("CS_NO_U2B_SPLIT", "1")
    monkeypatch.setenv("CS_NO_TOP32", "1")
    f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], p0, tile=128, inv_pipeline=1)
    g0, st0, mu0 = f.eval()
    f.close()
    assert np.array_equal(g1, g0) and np.array_equal(st1, st0) and mu1 == mu0
    g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
    assert cs.rel(g1, o.pack_params(g), 1e-9) < cs.TOL_GRAD and abs(st1[1] - st[1]) < cs.TOL_LML * abs(st[1])


def test_concurrent_fits_match_sequential(emulib):
    Ps = [cs.make_problem("GP", 90, 3, replicate=r) for r in range(3)]
    fits = [_lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), chunk=2) for P in Ps]
    res = _lib.run_many(fits, 7, 1e-12)
    for P, f, (it, tr) in zip(Ps, fits, res):
        ref = cs.oracle_fit(P, "Nadam", 7, 1e-12, 0.01)
        assert it == ref.iters and cs.rel(tr[1, 1:], ref.stats[1, 1:]) < 1e-10
        assert cs.rel(f.get_params(), o.pack_params(ref.Kernel.parameters), 1e-9) < 1e-8
        f.close()


@pytest.mark.parametrize("n,p,tile", [(150, 3, 64), (330, 4, 64), (300, 3, 128), (520, 2, 128)])
def test_pipelined_inverse(emulib, n, p, tile):
    P = cs.make_problem("GP", n, p, weights=True)
    g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
    f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=tile, inv_pipeline=1)
    g2, st2, mu2 = f.eval()
    g3, st3, _ = f.eval()
    f.close()
    assert np.array_equal(g2, g3)
    assert cs.rel(g2, o.pack_params(g), 1e-9) < cs.TOL_GRAD and abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1])
    assert abs(st2[0] - st[0]) < 1e-8 * abs(st[0]) and abs(mu2 - mu) < 1e-9 * abs(mu)


@pytest.mark.parametrize("kind,n,p,B", [("GP", 90, 3, 3), ("TP", 70, 2, 2), ("GP", 150, 4, 4)])
def test_batched_handle_matches_individual_fits(emulib, kind, n, p, B):
    """cs_fit_create_batch: every launch serves all members (blockIdx.z); members differ in data, start
    and stopping iteration.  Each member must equal the oracle fit of its own problem."""
    Ps = [cs.make_problem(kind, n, p, replicate=r, draws=(0.4 + 0.1 * r, 0.6 - 0.05 * r)) for r in range(B)]
    members = [(P["y"], P["X"], P["z"], P["w"] if r % 2 else None, o.pack_params(P["par"])) for r, P in enumerate(Ps)]
    fb = _lib.FitBatch(emulib, cs.kcode(kind), members, chunk=3)
    assert emulib.dll.cs_fit_batch_size(fb.h) == B
    fb.eval_async(); fb.sync()
    for b, P in enumerate(Ps):
        g, st, mu = o.eval_lml_grad(kind, P["y"], P["X"], P["z"], P["w"], P["par"])
        g2, st2, mu2 = fb.select(b).fetch()
        assert cs.rel(g2, o.pack_params(g), 1e-9) < cs.TOL_GRAD and abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1])
        assert abs(mu2 - mu) < 1e-9 * abs(mu)
    tol = 0.5  # loose: members stop at different iterations
    res = fb.select(0).run(14, tol)
    its = []
    for b, (P, (it, tr)) in enumerate(zip(Ps, res)):
        ref = cs.oracle_fit(P, "Nadam", 14, tol, 0.01)
        its.append(it)
        assert it == ref.iters, (b, it, ref.iters)
        assert cs.rel(tr[1, 1:], ref.stats[1, 1:]) < 1e-9
        assert cs.rel(fb.select(b).get_params(), o.pack_params(ref.Kernel.parameters), 1e-9) < 1e-8
        X2 = P["X"][:11] * 0.9
        t2 = fb.treat(X2, cov_jitter=1e-6 if kind == "TP" else 0.0)
        tr_o = ref.Kernel.predict_treat(P["y"], P["X"], P["z"], X2)
        assert cs.relm(t2["map"], tr_o["map"]) < cs.TOL_MAP
    fb.close()


@pytest.mark.parametrize("n,tile,pipe", [(600, 64, 1), (700, 64, 0), (1200, 128, 1)])
def test_graded_panel_widths(emulib, monkeypatch, n, tile, pipe):
    """Plan v2 with a narrow first panel and narrow last panels (the n = 8192 layout), forced at small N."""
    monkeypatch.setenv("CS_GRADED", "2")
    assert emulib.dll.cs_debug_check_plan(n, tile, 64, 1) == 0
    P = cs.make_problem("GP", n, 3)
    g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
    f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=tile, inv_pipeline=pipe)
    g2, st2, mu2 = f.eval()
    g3, st3, _ = f.eval()
    f.close()
    assert np.array_equal(g2, g3)
    assert cs.rel(g2, o.pack_params(g), 1e-9) < cs.TOL_GRAD and abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1])
    assert abs(mu2 - mu) < 1e-9 * abs(mu)


def test_limits_and_failure_reporting(emulib):
    cs.check_limits(emulib)


@pytest.mark.parametrize("pipe", [0, 1])
def test_first_generation_plan_still_works(emulib, monkeypatch, pipe):
    """CS_PLAN=1 keeps the per-step plan (panel multiply + update per 128 columns, memset-initialised inverse) as an
    A/B reference for the v2 plan; it must stay race-free and numerically equivalent."""
    monkeypatch.setenv("CS_PLAN", "1")
    assert emulib.dll.cs_debug_check_plan(3000, 128, 64, 1) == 0
    P = cs.make_problem("GP", 330, 4, weights=True)
    g, st, mu = o.eval_lml_grad("GP", P["y"], P["X"], P["z"], P["w"], P["par"])
    f = _lib.Fit(emulib, 0, P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=64, inv_pipeline=pipe)
    g2, st2, mu2 = f.eval()
    g3, _, _ = f.eval()
    f.close()
    assert np.array_equal(g2, g3)
    assert cs.rel(g2, o.pack_params(g), 1e-9) < cs.TOL_GRAD and abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1])


def test_failing_member_stops_alone(emulib):
    cs.check_failing_member(emulib)


def test_kernel_handle_cache_follows_the_data(emulib):
    """the RefClass mirror caches its device handle; the cache must notice in-place edits of y / X / z, must not
    be fooled by a reused id(), and must not re-create the handle for a 1-D X on every call."""
    from causalstump_b200 import main as M, rcpp_exports as rx
    rx.use_library(emulib)
    try:
        P = cs.make_problem("GP", 30, 1, readme=True)
        y, X1d, z = P["y"].copy(), P["X"][:, 0].copy(), P["z"].copy()
        K = M.KernelClass_GP_SE(1, np.ones(30))
        K.parameters = P["par"]
        s0 = K.get_train_stats(y, X1d, z)
        h0 = K._fit
        assert K.get_train_stats(y, X1d, z)[1] == s0[1] and K._fit is h0      # 1-D X: same handle
        y[3] += 0.5                                                           # in place: same id, new content
        s1 = K.get_train_stats(y, X1d, z)
        assert K._fit is not h0 and s1[1] != s0[1]
        want = o.eval_lml_grad("GP", y, X1d[:, None], z, np.ones(30), P["par"])[1][1]
        assert abs(s1[1] - want) < 1e-9 * abs(want)
        K.invalidate()
        assert K._fit is None
    finally:
        rx.use_library(None)


@pytest.mark.parametrize("kind,n,p", [("GP", 200, 25), ("TP", 131, 7), ("GP", 70, 32)])
def test_packed_tiles_and_cooperative_staging_agree(emulib, monkeypatch, kind, n, p):
    """The tensor-path Gram / gradient kernels read either the packed tiles written by pack_x_tiles_kernel (covariates, z, w,
    the scaled norms and the per-evaluation constants; on the GPU one bulk copy per operand tile, here plain copies) or stage
    the same layout cooperatively (CS_NO_TMA=1).  Same arithmetic on both routes: identical results, and both match the oracle."""
    P = cs.make_problem(kind, n, p, weights=True)
    g, st, mu = o.eval_lml_grad(kind, P["y"], P["X"], P["z"], P["w"], P["par"])
    out = []
    for no_tma in (False, True):
        if no_tma:
            monkeypatch.setenv("CS_NO_TMA", "1")
        else:
            monkeypatch.delenv("CS_NO_TMA", raising=False)
        f = _lib.Fit(emulib, cs.kcode(kind), P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]))
        try:
            out.append(f.eval())
        finally:
            f.close()
    (g1, st1, mu1), (g2, st2, mu2) = out
    assert np.array_equal(g1, g2) and np.array_equal(st1, st2) and mu1 == mu2
    gref = o.pack_params(g)   # entries that are ~0 by cancellation are compared against the largest one (as cs.check_eval)
    gerr = float(np.max(np.abs(g1 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))))
    assert gerr < cs.TOL_GRAD and abs(st1[1] - st[1]) < cs.TOL_LML * abs(st[1])
