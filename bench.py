#!/usr/bin/env python
"""bench.py -- fp64 LML+gradient evals/sec at n=8192, p=25 (BASELINE.json metric i) plus
Hill-sim replicate fits/hour (metric ii) as an extra object.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one LML+gradient evaluation of the GP on one Hill-(2011)-shaped synthetic
batch: Gram -> Cholesky -> inverse -> fused trace gradients -> LML / stats
-> closed-form mu.  `value` is device-resident throughput (inputs already in HBM);
`e2e` goes through the C ABI with host buffers (pinned-staged H2D of y, X, z, w and the
parameters, D2H of gradient + stats every step).  N > 1 runs one replica per GPU
("replicas only", no collective on the data path).

--impl reference times the reference's OWN C++ (oracle/_ref: src/kernel_GP_SE_cpp.cpp compiled unmodified against the
header shim of oracle/shim, arma::inv_sympd / log_det through the host LAPACK; kind "reference") on the host cores: ONE
full n=8192, p=25 evaluation, chained as para_update chains the routines.  R and RcppArmadillo are not installed in this
image; where oracle/_ref is absent the NumPy oracle port stands in (kind "port").

N > 1 additionally reports the two sharded configurations of BASELINE.json: `config4` (100 replicates x 4 restarts = 400
fits SPLIT over the N GPUs, no collective) and `config5` (one n=32768 evaluation on a P x Q grid, 2-D block-cyclic
Cholesky with NCCL panel broadcasts).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_METRIC, P_METRIC = 8192, 25
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed
# `ncu --set full` capture (per launch; None until a capture of the current kernel exists)
XTX_TRAFFIC_BYTES = 4.555720e9 + 486.406912e6
XTX_TRAFFIC_SOURCE = ("profiles/r02_ncu_xtx.txt: ncu --set full of this kernel in `bench.py --inv-pipeline 0` "
                      "(dram__bytes_read.sum 4.556 GB + dram__bytes_write.sum 486.4 MB per launch, final round-2 build)")
METRIC = "fp64 LML+gradient evals/sec at n=8192,p=25"
UNIT = "evals/s"
# time-weighted DMMA sub-pipe activity of the K = 512 bulk launches of the timed plan (ncu --set full over every GEMM launch
# of one evaluation, plain priority streams -- ncu cannot attach to green-context streams); updated with each capture
DMMA_ACTIVE_K512 = {"pct": 88.0, "all_gemm_launches_pct": 78.2,
                    "source": "profiles/r02_gemm_launches.md (71 launches with >= 592 CTA tiles, time-weighted; 78.2 % over all 460 GEMM launches; ncu pass taken before the last GEMM change of the round -- fixed per-thread copy offsets -- which lifted a K = 512 NT launch from 33.5 to 34.5 TFLOP/s)"}


def workload_config():
    """The `config` object: identical in both arms (ours / --impl reference) at every N."""
    return {"workload": "Hill-(2011)-shaped synthetic n=8192 p=25, GP (prior=FALSE), one LML+gradient evaluation per step",
            "n": N_METRIC, "p": P_METRIC, "l2": "working set 3 x 512 MiB per evaluation, larger than the 126 MB L2 (no flush needed)"}


def check_fracs(obj, path="line"):
    """No roofline fraction above 1.2 may be printed (a bracket that does not contain the work it claims)."""
    if isinstance(obj, dict):
        for k, v in obj.items():
            if k == "frac" and isinstance(v, (int, float)):
                assert v <= 1.2, "%s.frac = %.3f: the timed bracket does not contain the work" % (path, v)
            check_fracs(v, path + "." + str(k))


def make_inputs(n, p, replicate, kind="GP"):
    """Hill-shaped synthetic data, reference normalisation and reference initial
    hyper-parameters (R/kernel_GP_SE_class.R:13-19).  Host prep, not on the timed path."""
    from causalstump_b200 import hill
    from causalstump_b200.main import KernelClass_GP_SE, KernelClass_TP_SE, norm_variables
    from causalstump_b200 import rcpp_exports as rx
    d = hill.hill_surface_b(n=n, p=p, replicate=replicate)
    nr = norm_variables(d["y"], d["X"])
    y, X, z = np.ascontiguousarray(nr["y"]), np.asfortranarray(nr["X"]), np.ascontiguousarray(d["z"])
    draws = hill.restart_draws(replicate, 0, prior=(kind == "TP"))
    K = KernelClass_GP_SE(p, np.ones(n)) if kind == "GP" else KernelClass_TP_SE(p, np.ones(n))
    if kind == "GP":
        K.parainit(y, draws)
    else:
        K.parainit(y, 2000, draws)
    return dict(y=y, X=X, z=z, w=np.ones(n), params=rx.pack(K.parameters), raw=d, moments=nr["moments"], par=K.parameters)


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU side: the oracle on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------
def cpu_eval_seconds(inp, budget_s):
    """Reference-faithful CPU cost of ONE n=8192 evaluation, from a bounded sample:
    the O(p n^2) element-wise passes are timed on a few of the p dimensions at full n
    and scaled by p/dims; the LAPACK part (dpotrf + dpotri + mirror, dgetrf of the
    inverse, mat-vecs) is timed at n_s <= n and scaled by (n/n_s)^3.  Returns
    (seconds, description)."""
    from oracle import causalstump_oracle as o
    from scipy.linalg import lapack
    y, X, z, w, par = inp["y"], inp["X"], inp["z"], inp["w"], inp["par"]
    n, p = X.shape
    t_all0 = time.perf_counter()
    # --- Gram passes (kernmat_GP_SE_cpp :105-134) on dg dims
    dg = 2 if budget_s < 20 else 3
    sub = dict(par); sub["Lm"] = par["Lm"][:dg]; sub["La"] = par["La"][:dg]
    t0 = time.perf_counter()
    Kl = o.kernmat_GP_SE_cpp(X[:, :dg], X[:, :dg], z, z, sub)
    t_gram_dims = time.perf_counter() - t0
    # split fixed part (final exp pass) from per-dim part with a second, 1-dim measurement
    sub1 = dict(par); sub1["Lm"] = par["Lm"][:1]; sub1["La"] = par["La"][:1]
    t0 = time.perf_counter()
    o.kernmat_GP_SE_cpp(X[:, :1], X[:, :1], z, z, sub1)
    t_gram_1 = time.perf_counter() - t0
    per_dim = max((t_gram_dims - t_gram_1) / (dg - 1), 0.0)
    t_gram = t_gram_1 + per_dim * (p - 1)
    # --- dense LAPACK part at n_s
    ns = n
    while ns > 1024 and 6.0 * (ns / 4096.0) ** 3 > 0.45 * budget_s:
        ns //= 2
    A = np.array(Kl["Kmat"][:ns, :ns], order="F")
    A[np.diag_indices(ns)] += 1.0
    t0 = time.perf_counter()
    inv = o.inv_sympd(A)
    o.log_det(inv)
    t_dense = (time.perf_counter() - t0) * (n / ns) ** 3
    # --- gradient passes (evid_scale_gradients :28-41) on dgr dims at full n
    dgr = 1
    T = Kl["Km"]  # any dense n x n stand-in for tmpK: the passes are data-independent
    t0 = time.perf_counter()
    o.evid_scale_gradients(X[:, :dgr], T, Kl["Km"], Kl["Ka"], dict(Lm=par["Lm"][:dgr], La=par["La"][:dgr]))
    t_grad = (time.perf_counter() - t0) * (p / dgr)
    # --- the remaining n^2 passes of grad_GP_SE_cpp (:158-198): alpha, outer product, 3 traces, K alpha
    ybar = y - par["mu"]
    t0 = time.perf_counter()
    alpha = T @ ybar
    tmpK = T - np.outer(alpha, alpha)
    o.evid_grad(tmpK, Kl["Km"]); o.evid_grad(tmpK, Kl["Ka"]); o.evid_grad(tmpK, np.diag(1.0 / w))
    (Kl["Kmat"] @ alpha).sum()
    t_rest = time.perf_counter() - t0
    total = t_gram + t_dense + t_grad + t_rest
    desc = ("oracle (NumPy + SciPy-OpenBLAS restatement of the reference), n=%d p=%d: Gram passes timed on %d of %d dims, "
            "lengthscale-gradient passes on %d of %d dims (scaled by p/dims), inv_sympd+log_det timed at n_s=%d "
            "(scaled by (n/n_s)^3), remaining n^2 passes in full; %.1f s of CPU work"
            % (n, p, dg, p, dgr, p, ns, time.perf_counter() - t_all0))
    return total, desc


def ref_eval_seconds(inp, budget_s):
    """Cost of ONE n=8192 evaluation by the reference's OWN C++ (oracle/_ref/libcausalstump_ref.so: its three source
    files compiled unmodified; inv_sympd / log_det through the host LAPACK, as Armadillo does), from a bounded sample.
    The routines are called the way para_update chains them (R/kernel_GP_SE_class.R:45-60) on the first n_s rows:
    kernmat and grad with 1 and with 2 covariate columns (per-dimension cost = the difference, scaled to p),
    invkernel and log_det (n^3) scaled by (n/n_s)^3, everything else by (n/n_s)^2.  Returns (seconds, description)
    or None when the compiled reference is not present."""
    from oracle import ref_binding as rb
    if not rb.available():
        return None
    ref = rb.RefLib()
    lap = ref.use_host_lapack()
    y, X, z, w, par = inp["y"], inp["X"], inp["z"], inp["w"], inp["par"]
    n, p = X.shape
    ns = n if budget_s >= 45 else (n // 2 if budget_s >= 8 else n // 4)
    r2, r3 = (n / ns) ** 2, (n / ns) ** 3
    ys, zs, ws = y[:ns].copy(), z[:ns].copy(), w[:ns].copy()
    t_all0 = time.perf_counter()

    def kern(d):
        Xd = np.asfortranarray(X[:ns, :d])
        t0 = time.perf_counter()
        out = ref.kernmat(0, Xd, Xd, zs, zs, par["Lm"][:d], par["La"][:d], par["lambdam"], par["lambdaa"])
        return time.perf_counter() - t0, out

    def flat(d):
        return np.concatenate([[par["sigma"], par["lambdam"], par["lambdaa"]], par["Lm"][:d], par["La"][:d], [par["mu"]]])

    k1, _ = kern(1)
    k2, (K, Km, Ka) = kern(2)
    k_dim = max(k2 - k1, 0.0)
    t_kern = (max(k1 - k_dim, 0.0) + p * k_dim) * r2
    t0 = time.perf_counter()
    invK = ref.invkernel(zs, ws, K, par["sigma"])
    t_inv = (time.perf_counter() - t0) * r3
    t0 = time.perf_counter()
    ref.log_det(invK)
    lu = time.perf_counter() - t0
    g = []
    for d in (1, 2):
        Xd = np.asfortranarray(X[:ns, :d])
        t0 = time.perf_counter()
        ref.grad(0, ys, Xd, zs, ws, K, Km, Ka, invK, flat(d))
        g.append(time.perf_counter() - t0)
    g_dim = max(g[1] - g[0], 0.0)
    t_grad = (max(g[0] - g_dim - lu, 0.0) + p * g_dim) * r2 + lu * r3
    t0 = time.perf_counter()
    ref.mu_solution(ys, zs, invK)
    t_mu = (time.perf_counter() - t0) * r2
    total = t_kern + t_inv + t_grad + t_mu
    desc = ("the reference's own C++ (src/kernel_GP_SE_cpp.cpp compiled unmodified against oracle/shim -- an eager stand-in for Armadillo: "
            "the reference's loops run verbatim, Armadillo's expression templates do not; LAPACK = %s) on the "
            "first n_s=%d rows of the n=%d p=%d workload: kernmat and grad timed with 1 and 2 covariate columns (per-column cost "
            "scaled to p=%d), n^2 parts scaled by (n/n_s)^2, invkernel and log_det by (n/n_s)^3; components at n: kernmat %.1f s, "
            "invkernel %.1f s, grad %.1f s, mu %.2f s; %.1f s of CPU work"
            % (lap, ns, n, p, p, t_kern, t_inv, t_grad, t_mu, time.perf_counter() - t_all0))
    return total, desc


def host_env():
    """OMP / BLAS thread settings and the BLAS build, for the cpu_baseline record."""
    info = {"OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS"), "OPENBLAS_NUM_THREADS": os.environ.get("OPENBLAS_NUM_THREADS"),
            "nproc": os.cpu_count()}
    try:
        from threadpoolctl import threadpool_info
        for i in threadpool_info():
            if i.get("user_api") == "blas":
                info["blas"] = "%s %s (%s, %d threads)" % (i.get("internal_api"), i.get("version"), i.get("architecture"), i.get("num_threads", 0))
                break
    except Exception:
        pass
    return info


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        ths = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        if ths:
            return int(max(ths))
    except Exception:
        pass
    return os.cpu_count() or 1


def ref_available():
    from oracle import ref_binding as rb
    return rb.available()


def ref_full_eval(inp, threads):
    """ONE full evaluation by the reference's own compiled routines, chained as para_update chains them
    (R/kernel_GP_SE_class.R:45-60): kernmat -> invkernel -> grad -> mu_solution.  Returns per-routine seconds, the LML,
    the LAPACK description, and the two LAPACK-bound parts (inv_sympd inside invkernel, log_det inside grad) re-timed alone
    at `threads` and at 2 BLAS threads (OMP_NUM_THREADS=2 is the reference's documented setting, README.md:13)."""
    from oracle import ref_binding as rb
    from threadpoolctl import threadpool_limits
    ref = rb.RefLib()
    t = {"lapack": ref.use_host_lapack()}
    y, X, z, w, par = inp["y"], inp["X"], inp["z"], inp["w"], inp["par"]
    flat = np.asarray(inp["params"], dtype=np.float64)
    with threadpool_limits(limits=threads, user_api="blas"):
        t0 = time.perf_counter()
        K, Km, Ka = ref.kernmat(0, X, X, z, z, par["Lm"], par["La"], par["lambdam"], par["lambdaa"])
        t["kernmat"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        invK = ref.invkernel(z, w, K, par["sigma"])
        t["invkernel"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        g, st = ref.grad(0, y, X, z, w, K, Km, Ka, invK, flat)
        t["grad"] = time.perf_counter() - t0
        del Km, Ka
        t0 = time.perf_counter()
        ref.mu_solution(y, z, invK)
        t["mu_solution"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref.log_det(invK)
        t["log_det_alone"] = time.perf_counter() - t0
    with threadpool_limits(limits=2, user_api="blas"):
        t0 = time.perf_counter()
        ref.invkernel(z, w, K, par["sigma"])
        t["invkernel_2threads"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref.log_det(invK)
        t["log_det_2threads"] = time.perf_counter() - t0
    t["lml"] = float(st[1])
    return t


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    ncores = os.cpu_count() or 1
    try:  # torchrun exports OMP_NUM_THREADS=1 to every rank: give the reference arm all host threads back
        import scipy.linalg  # noqa: F401  (load SciPy's OpenBLAS before limits are applied)
        from oracle import causalstump_oracle  # noqa: F401
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncores, user_api="blas")
    except Exception:
        pass
    inp = make_inputs(N_METRIC, P_METRIC, 0)
    extra = {}
    if ref_available():
        # The whole budget of this arm goes into ONE honest full-size evaluation (about 3 minutes on 16 host cores);
        # --steps / --warmup are not multiplied into it: the line says how many steps were really run.
        kind = "reference"
        t = ref_full_eval(inp, ncores)
        secs = t["kernmat"] + t["invkernel"] + t["grad"] + t["mu_solution"]
        desc = ("ONE full evaluation at n=%d p=%d by the reference's own C++ (src/kernel_GP_SE_cpp.cpp compiled unmodified, -O2) chained "
                "as para_update chains it: kernmat %.1f s, invkernel %.1f s, grad %.1f s (of which log_det %.1f s), mu_solution %.2f s; "
                "arma::inv_sympd / log_det = host LAPACK dpotrf+dpotri / dgetrf (%s, %d threads); every other Armadillo operation is "
                "the EAGER stand-in of oracle/shim/RcppArmadillo.h (the reference's own loops run verbatim, Armadillo's expression "
                "templates -- which fuse some temporaries of the O(p n^2) passes -- do not), single-threaded like Armadillo's"
                % (N_METRIC, P_METRIC, t["kernmat"], t["invkernel"], t["grad"], t["log_det_alone"], t["mu_solution"], t["lapack"], ncores))
        secs2 = secs - t["invkernel"] - t["log_det_alone"] + t["invkernel_2threads"] + t["log_det_2threads"]
        extra["omp_num_threads_2"] = {"value": 1.0 / secs2, "unit": UNIT, "seconds": secs2,
                                      "how": "the full run above with invkernel (%.1f s) and log_det (%.1f s) re-timed at 2 BLAS threads; "
                                             "the element-wise passes are single-threaded at any setting"
                                             % (t["invkernel_2threads"], t["log_det_2threads"])}
        m = ref_eval_seconds(inp, 8.0)  # the round-1 bounded-sample model, kept beside the measurement
        if m is not None:
            extra["model"] = {"value": 1.0 / m[0], "unit": UNIT, "seconds": m[0], "ratio_to_measured": m[0] / secs, "sample": m[1]}
        extra["lml"] = t["lml"]
        steps_run, sample_kind = 1, "full"
    else:  # oracle/_ref was not built (no /root/reference at build time): the oracle port stands in, bounded sample
        kind = "port"
        secs, desc = cpu_eval_seconds(inp, 30.0)
        steps_run, sample_kind = 1, "bounded sample"
    ms = 1e3 * secs
    value = 1.0 / secs
    ps, pdesc = cpu_eval_seconds(inp, 12.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_run, "warmup": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup, "step_is": sample_kind,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(), "parallelism": "rank 0 only: one host process, all host threads",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": desc, "host": host_env(),
                         "oracle_port": {"value": 1.0 / ps, "unit": UNIT, "sample": pdesc}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"].update(extra)
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 burst (best of 5) -- the FP64 denominator MEASURED_PEAKS.json lacks."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_units(units, batch, host_sim):
    """The replicate loop: on the device end to end (cs_sim_run: generator, split, fits, predict / treatment, metrics) or,
    with --host-sim, with the generator / split / metrics in host NumPy around the device fits (the round-1 driver)."""
    from causalstump_b200 import replicates as rp
    return rp.hill_units_concurrent(units, batch=batch) if host_sim else rp.hill_units_device(units, batch=batch)


def measure_fits(lib, rank, world, nfit, batch=32, host_sim=False):
    """Metric ii: one 'fit' = CausalStump() on a 672 x 25 training split to the reference's
    stopping rule (maxiter=3000, tol=1e-1, lr=.01, Nadam; test/sec5-3_Hill2011.R:344-345)
    + predict on 747 + treatment on 672 and 75.  The nfit fits of this rank (replicates x
    4 restarts, config 4) are optimised concurrently (cs_fit_run_many).
    Returns (seconds, fits, iterations, mean PEHE of the restart winners)."""
    from causalstump_b200 import replicates as rp
    units = [(rank * ((nfit + 3) // 4) + k // 4, k % 4) for k in range(nfit)]
    run_units(units[:2], batch, host_sim)  # warm-up: library load, graph capture, allocator
    t0 = time.perf_counter()
    res = run_units(units, batch, host_sim)
    secs = time.perf_counter() - t0
    win = rp.select_winners(res)
    return secs, nfit, [r["iters"] for r in res], float(np.mean([w["pehe_test"] for w in win]))


GRIDS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}


def measure_single_fit_configs():
    """Configs 1-3 of BASELINE.json through the PUBLIC API (the calls a user of the reference makes): CausalStump() with the
    reference's default settings (maxiter 5000, tol 1e-4, lr .01, Nadam) to convergence, then predict() and treatment();
    wall clock, host buffers in and out.  config1: README example (n=120, p=1, GP); config2: Hill-shaped n=747, p=25, GP;
    config3: the same data, Student-t process (prior=TRUE, nu=200) incl. posterior sampling of the intervals."""
    import causalstump_b200 as cb
    from causalstump_b200 import hill
    out = {}
    for name, d, kw in (("config1", hill.readme_example(120), dict(prior=False)),
                        ("config2", hill.hill_surface_b(n=747, p=P_METRIC, replicate=3), dict(prior=False)),
                        ("config3", hill.hill_surface_b(n=747, p=P_METRIC, replicate=3), dict(prior=True, nu=200))):
        cb.CausalStump(d["y"], d["X"], d["z"], maxiter=3, init_draws=(0.5, 0.4), verbose=False, **kw)  # warm-up (graph capture)
        t0 = time.perf_counter()
        fit = cb.CausalStump(d["y"], d["X"], d["z"], init_draws=(0.5, 0.4), verbose=False, **kw)
        t_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        pr = cb.predict(fit)
        tr = cb.treatment(fit)
        t_pt = time.perf_counter() - t0
        pehe = float(np.sqrt(np.mean((d["tau"] - tr["map"]) ** 2)))
        out[name] = {"n": int(len(d["y"])), "p": int(np.atleast_2d(d["X"].T).shape[0]), "prior": bool(kw["prior"]), "iterations": int(fit.iters),
                     "fit_seconds": t_fit, "ms_per_iteration": 1e3 * t_fit / max(int(fit.iters), 1), "predict_plus_treatment_seconds": t_pt,
                     "lml": float(fit.stats[1, -1]), "pehe_train": pehe, "ate": float(tr["ate"]),
                     "finite": bool(np.isfinite(pr["map"]).all() and np.isfinite(tr["map"]).all())}
    out["what"] = ("BASELINE.json configs 1-3 through the public API (CausalStump / predict / treatment, reference defaults: maxiter 5000, "
                   "tol 1e-4, lr .01, Nadam), wall clock incl. host<->device copies; the whole optimisation loop is one device-side loop")
    return out


def measure_stateless_boundary(sizes=((747, 3), (8192, 1))):
    """What a user gets who swaps the library in but leaves the R package untouched (INTEGRATION.md step 2 without step 3):
    the RefClass methods keep issuing the 8 stateless `.Call`s of one GP iteration (R/kernel_GP_SE_class.R:45-60: kernmat,
    invkernel, grad, Nadam, mu_solution, then get_train_stats = kernmat, invkernel, stats), each shipping its n x n matrices
    over PCIe in both directions.  Timed through the host mirrors of those routines (causalstump_b200/rcpp_exports.py), wall
    clock, host buffers; next to it the same iteration on the device handle (the fast path CausalStump() uses)."""
    from causalstump_b200 import _lib, rcpp_exports as rx
    from causalstump_b200.main import list2zero
    out = {"what": "one GP optimizer iteration through the 8 stateless .Call mirrors (n x n matrices cross PCIe on every call) vs the "
                   "same iteration on the device-resident handle; seconds per iteration, wall clock", "sizes": []}
    for n, reps in sizes:
        inp = make_inputs(n, P_METRIC, 1)
        y, X, z, w, par = inp["y"], inp["X"], inp["z"], inp["w"], inp["par"]
        m, v = list2zero(par), list2zero(par)

        def iteration(it, par, m, v):
            Kl = rx.kernmat_GP_SE_cpp(X, X, z, z, par)
            inv = rx.invkernel_cpp(z, w, Kl["Kmat"], par)
            gl = rx.grad_GP_SE_cpp(y, X, z, w, Kl["Kmat"], Kl["Km"], Kl["Ka"], inv, par)
            o = rx.Nadam_cpp(it, 0.01, 0.9, 0.999, 1e-8, m, v, gl["gradients"], par)
            par = o["parameters"]
            par["mu"] = rx.mu_solution_cpp(y, z, inv, par)
            Kl = rx.kernmat_GP_SE_cpp(X, X, z, z, par)          # get_train_stats (:58)
            inv = rx.invkernel_cpp(z, w, Kl["Kmat"], par)
            st = rx.stats_GP_SE(y, Kl["Kmat"], inv, par)
            return par, o["m"], o["v"], st

        par1, m1, v1, _ = iteration(1, par, m, v)  # warm-up
        t0 = time.perf_counter()
        for it in range(reps):
            par1, m1, v1, st = iteration(it + 2, par1, m1, v1)
        sec_stateless = (time.perf_counter() - t0) / reps
        f = _lib.Fit(rx.lib(), _lib.CS_GP, y, X, z, w, inp["params"])
        f.run(3, 0.0)
        iters = 20 if n < 4096 else 5
        t0 = time.perf_counter()
        f.run(iters, 0.0)
        sec_handle = (time.perf_counter() - t0) / (iters + 1)   # + the final get_train_stats evaluation
        f.close()
        pcie = 8.0 * n * n * (3 + 1 + 1 + 4 + 1 + 3 + 1 + 1 + 2)   # matrices out/in per iteration: 3, 1+1, 4, 1, 3, 1+1, 2
        out["sizes"].append({"n": n, "p": P_METRIC, "stateless_s_per_iteration": sec_stateless, "handle_s_per_iteration": sec_handle,
                             "ratio": sec_stateless / sec_handle, "pcie_bytes_per_iteration": pcie, "lml_after": float(st[1])})
    return out


def measure_config4(lib, rank, world, dist, barrier, max_over_ranks, batch, host_sim=False):
    """Config 4 AS STATED (test/sec5-3_Hill2011.R:611-616): 100 replicates x 4 restarts = 400 independent fits SPLIT over the
    N GPUs (static round-robin of (replicate, restart) units, no collective; the result rows are gathered on rank 0 and the
    restart with the highest final LML wins per replicate).  Strong scaling: 400 / N fits per GPU."""
    from causalstump_b200 import replicates as rp
    units = rp.shard(rp.unit_list(100, 4), rank, world)
    # few units per GPU: smaller batched handles keep more handles in flight (measured on one B200, device driver:
    # 50 units 1.54 s at batch 8 vs 1.68 s at 32; 100 units 2.46 s at 16 vs 2.67 s at 32; 200 units: 32 as good as 8)
    batch = batch if len(units) > 128 else (16 if len(units) > 64 else 8)
    run_units(units[:2], batch, host_sim)  # warm-up: graph capture, allocator
    barrier()
    t0 = time.perf_counter()
    res = run_units(units, batch, host_sim)
    secs_local = time.perf_counter() - t0
    tim = dict(rp.LAST_TIMING)
    secs = max_over_ranks(secs_local)
    rows = [(r["replicate"], r["restart"], r["lml"], r["iters"], r["pehe_test"], r["failed"]) for r in res]
    util = [(tim["iterations"], tim["iterations"] / max(tim["slot_utilisation"], 1e-12), secs_local)]
    if dist is not None:
        allr, allu = [None] * world, [None] * world
        dist.all_gather_object(allr, rows)
        dist.all_gather_object(allu, util)
        rows = [x for part in allr for x in part]
        util = [x for part in allu for x in part]
    if rank != 0:
        return None
    assert sorted((a, b) for a, b, *_ in rows) == rp.unit_list(100, 4), "every (replicate, restart) unit exactly once"
    win = rp.select_winners([dict(replicate=a, restart=b, lml=c, pehe_test=e) for a, b, c, d, e, f in rows if not f])
    its = [d for a, b, c, d, e, f in rows]
    return {"what": "config 4 as stated: 100 Hill surface-B replicates x 4 restarts = 400 GP fits (672 x 25 training split, reference stopping "
                    "rule maxiter 3000 / tol 1e-1 / lr .01 Nadam, + predict(747) + treatment(672, 75) + the 22 metrics) split over "
                    "%d GPU(s), %d fits per GPU in batched handles of %d; %s; wall time of the whole loop incl. handle creation, "
                    "max over ranks" % (world, len(units), batch, "generator / split / metrics in host NumPy (--host-sim)" if host_sim else
                                        "simulation driver on the device (cs_sim_run: Philox response-surface generator, split, "
                                        "normalisation, fits, predict / treatment and the metric block never leave the GPU)"),
            "fits": len(rows), "fits_per_gpu": len(units), "seconds": secs, "value": len(rows) / secs * 3600.0, "unit": "fits/hour",
            "scaling": "strong", "iterations_mean": float(np.mean(its)), "iterations_max": int(np.max(its)),
            "slot_utilisation": float(sum(u[0] for u in util) / max(sum(u[1] for u in util), 1e-12)),
            "seconds_per_rank": [round(u[2], 3) for u in util], "failed_units": int(sum(1 for r in rows if r[5])),
            "pehe_test_winners_mean": float(np.mean([w["pehe_test"] for w in win])), "phases_s_rank0": {k: round(v, 3) for k, v in tim.items()}}


def measure_config5(lib, torch, dev, rank, world, dist, barrier, steps=2, n=32768, parity_n=1500):
    """Config 5: ONE large evaluation (n=32768, p=25) on a P x Q grid of ranks: tiles 2-D block-cyclic, every rank builds its
    own Gram tiles, right-looking Cholesky / inverse / X^T X sweeps with NCCL panel broadcasts on row / column communicators
    (csrc/dist.cuh).  Parity of the same code path against the oracle at n=1500 first."""
    from causalstump_b200 import _lib
    P, Q = GRIDS[world]

    def make(nn):
        inp = make_inputs(nn, P_METRIC, 0)
        ids = [lib.dist_unique_id() if rank == 0 else None]
        if dist is not None:
            dist.broadcast_object_list(ids, src=0)
        return inp, _lib.DistFit(lib, 0, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], rank, world, P, Q, ids[0])

    out = {"what": "config 5: one LML+gradient evaluation at n=%d, p=%d, 2-D block-cyclic (128 x 128 tiles) over a %dx%d grid, NCCL panel "
                   "broadcasts; time = max over ranks of the device time of one evaluation" % (n, P_METRIC, P, Q),
           "n": n, "p": P_METRIC, "n_gpus": world, "grid": [P, Q]}
    inp, f = make(parity_n)
    g, st, mu, _ = f.eval()
    f.close()
    gathered = [(g.tolist(), st.tolist())]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, (g.tolist(), st.tolist()))
    if rank == 0:
        from oracle import causalstump_oracle as o   # the checker, outside every timed region
        go, sto, muo = o.eval_lml_grad("GP", inp["y"], inp["X"], inp["z"], inp["w"], inp["par"])
        gref = o.pack_params(go)
        out["parity"] = {"n": parity_n, "vs": "oracle",
                         "grad_rel": float(np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref))))),
                         "lml_rel": float(abs(st[1] - sto[1]) / abs(sto[1])), "mu_rel": float(abs(mu - muo) / abs(muo)),
                         "ranks_agree": bool(all(np.allclose(gathered[0][0], x[0], rtol=1e-12, atol=0) for x in gathered))}
    inp, f = make(n)
    f.eval()  # warm-up
    barrier()
    ms = []
    for _ in range(steps):
        g, st, mu, m = f.eval()
        ms.append(m)
    barrier()
    t = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    f.close()
    single = None
    if world == 1:
        # the same evaluation on the single-GPU engine (cs_fit: 512-column panels, SM partition, pipelined inverse) -- what a
        # user with one GPU would run; the distributed engine above still advances in 128-column steps
        inp1 = make_inputs(n, P_METRIC, 0)
        f1 = _lib.Fit(lib, _lib.CS_GP, inp1["y"], inp1["X"], inp1["z"], inp1["w"], inp1["params"], inv_pipeline=1)
        f1.eval_async(); f1.sync()
        f1.event_record(6)
        for _ in range(steps):
            f1.eval_async()
        f1.event_record(7)
        f1.sync()
        ms1 = f1.event_elapsed_ms(6, 7) / steps
        _, st1, _ = f1.fetch()
        f1.close()
        single = {"ms_per_eval": ms1, "tflops_n3": float(n) ** 3 / (ms1 * 1e-3) / 1e12, "lml": float(st1[1])}
    if rank != 0:
        return None
    ms_eval = float(t[0])
    if single:
        out["single_gpu_engine"] = single
    out.update({"ms_per_eval": ms_eval, "evals_per_s": 1e3 / ms_eval, "tflops_n3": float(n) ** 3 / (ms_eval * 1e-3) / 1e12,
                "steps": steps, "lml": float(st[1]), "finite": bool(np.isfinite(g).all() and np.isfinite(st).all()),
                "nccl_ranks": world, "scaling": "strong"})
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    from causalstump_b200 import _lib
    lib = _lib.load()
    lib.require_device()
    lib.check(lib.dll.cs_set_device(local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    inp = make_inputs(N_METRIC, P_METRIC, rank)  # one replicate per rank (weak scaling)
    fit = _lib.Fit(lib, _lib.CS_GP, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], inv_pipeline=args.inv_pipeline, partition=args.partition)
    n, p = N_METRIC, P_METRIC
    K, W = args.steps, max(args.warmup, 0)

    # ---- device-resident throughput (`value`) ----
    # nvidia-smi needs ~100 ms to deliver its first row and the timed region of 10 steps is only 0.2 s: the sampler
    # runs from the warm-up evaluations to the end of the end-to-end loop (the same workload throughout)
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(W, 3)):
        fit.eval_async()
    fit.sync()
    barrier()
    l0 = lib.launch_count()
    fit.dense_timing(True)   # event pair around the dense phase of every timed evaluation (the `roofline` object)
    fit.event_record(0)
    for _ in range(K):
        fit.eval_async()
    fit.event_record(1)
    fit.sync()
    barrier()
    launches = lib.launch_count() - l0
    dense_ms_total, dense_evals = fit.dense_ms()
    fit.dense_timing(False)
    assert dense_evals == K
    ms_total = max_over_ranks(fit.event_elapsed_ms(0, 1))
    g, st, mu = fit.fetch()
    assert np.isfinite(st).all() and np.isfinite(g).all(), "non-finite evaluation"
    ms_step = ms_total / K
    value = world * K / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers (`e2e`) ----
    h2d = 8 * (n * (p + 3) + len(inp["params"]))
    d2h = 8 * (len(inp["params"]) + 2 + 9) + 4
    for _ in range(2):
        fit.set_data(inp["y"], inp["X"], inp["z"], inp["w"]); fit.eval(inp["params"])
    barrier()
    t0 = time.perf_counter()
    fit.event_record(2)
    for _ in range(K):
        fit.set_data(inp["y"], inp["X"], inp["z"], inp["w"])
        fit.eval(inp["params"])
    fit.event_record(3)
    fit.sync()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_s = max_over_ranks(max(e2e_wall, fit.event_elapsed_ms(2, 3) * 1e-3))
    e2e = {"value": world * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    clk = clocks.stop()

    # ---- per-phase device times and the roofline of the dominant kernel ----
    # `phases` comes from the plan that was timed above.  With the pipelined plan the three dense
    # phases overlap on eight streams (their total is reported under "potrf"), so the per-launch
    # duration of the dominant kernel is measured on a second handle that runs the same kernels
    # back to back on one stream (inv_pipeline=0): K^-1 = X^T X there is ONE launch.
    phases = None
    for _ in range(3):
        ph = fit.eval_profile()
        phases = ph if phases is None else {k: min(phases[k], ph[k]) for k in ph}
    seq_phases = phases
    if args.inv_pipeline:
        fit_seq = _lib.Fit(lib, _lib.CS_GP, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], inv_pipeline=0)
        seq_phases = None
        for _ in range(4):
            ph = fit_seq.eval_profile()
            seq_phases = ph if seq_phases is None else {k: min(seq_phases[k], ph[k]) for k in ph}
        fit_seq.close()
    line_extra = {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = measure_fp64_peak(torch, dev)
        npad = n
        f_xtx = npad ** 3 / 3.0
        xtx_tflops = f_xtx / (seq_phases["xtx"] * 1e-3) / 1e12
        dense_ms = dense_ms_total / K        # measured INSIDE the timed region, on the handle's stream
        dense_tflops = float(npad) ** 3 / (dense_ms * 1e-3) / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line_extra["roofline"] = {
            "kernel": "csg::gemm_kernel (DMMA.8x8x4) + leaf kernel: the dense path of the TIMED plan -- Cholesky, triangular inverse and "
                      "K^-1 = X^T X, %s; every launch is an instance of the GEMM kernel template or the leaf kernel" %
                      ("inverse pipelined behind the Cholesky on 8 streams (look-ahead of depth three, chain on its own SM partition)"
                       if args.inv_pipeline else "phases back to back"),
            "bound": "tensor", "achieved": dense_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": dense_tflops / fp64_peak, "traffic": None,
            "flops_per_launch": float(npad) ** 3, "ms_per_launch": dense_ms,
            "launch": "one launch = the dense phase of one evaluation (about 460 GEMM / leaf launches overlapping on 8 streams); "
                      "algorithmic flops n^3 (n^3/3 potrf + 2n^3/3 potri); duration = CUDA event pair on the handle's stream around "
                      "that phase, averaged over the %d timed steps" % K,
            "dmma_active_pct_k512_launches": DMMA_ACTIVE_K512,
            "peak_source": "cuBLAS DGEMM 8192^3 fp64 burst measured live in this run (MEASURED_PEAKS.json has no fp64 entry)",
        }
        line_extra["roofline_kernel_isolated"] = {
            "kernel": "csg::gemm_kernel<64,2,2,true,true,16,2> (K^-1 = X^T X as ONE launch of 8256 CTA tiles, K = 8192), timed alone on its "
                      "stream on a second handle running the sequential plan -- NOT a launch of the timed step, kept as the ceiling "
                      "the K = 512 slices of the timed plan are compared with",
            "bound": "tensor", "achieved": xtx_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": xtx_tflops / fp64_peak, "traffic": XTX_TRAFFIC_BYTES, "traffic_source": XTX_TRAFFIC_SOURCE,
            "flops_per_launch": f_xtx, "ms_per_launch": seq_phases["xtx"],
        }
        if args.inv_pipeline:
            line_extra["phases_ms_sequential_plan"] = seq_phases
        # Gram / trace-gradient kernels: FP64-issue-bound at p = 25 (19 / 14 flop per byte against a ridge
        # of ~6), so the bound is the DFMA issue rate; the HBM figure north_star asks for is reported beside it.
        # cs_fit_eval_profile builds the whole Gram matrix on the handle's stream (no split), so `gram` brackets all of it.
        b_gram = 8.0 * n * n / 2 + 8.0 * n * (p + 2)   # lower triangle written once
        b_grad = 8.0 * n * n / 2 + 8.0 * n * (p + 3)   # lower triangle of K^-1 read once
        dfma_peak = 33.0  # TFLOP/s, DFMA-only issue rate measured by tools/pipe_probe.cu (profiles/r01_s2_pipe_probe.txt)
        f_gram = (n * n / 2.0) * (6 * p + 6)
        f_grad = (n * n / 2.0) * (10 * p + 12)
        ew = {}
        for nm, fl, by in (("gram", f_gram, b_gram), ("grad", f_grad, b_grad)):
            ms_k = seq_phases[nm]
            tf = fl / (ms_k * 1e-3) / 1e12
            ew[nm] = {"bound": "fp64-issue", "achieved": tf, "peak": dfma_peak, "unit": "TFLOP/s", "frac": tf / dfma_peak, "ms": ms_k,
                      "hbm_gbs": by / (ms_k * 1e-3) / 1e9, "hbm_frac": by / (ms_k * 1e-3) / 1e9 / hbm_peak}
        ew["note"] = ("both kernels timed as ONE launch each on the sequential-plan handle (the Gram build is not split there); fp64 flops "
                      "exclude the two exp per entry; peak = DFMA-only issue rate of tools/pipe_probe.cu (vector and tensor FP64 share "
                      "one pipe on sm_100a); hbm_* = algorithmic bytes / time against MEASURED_PEAKS.json hbm_gbs")
        line_extra["roofline_elementwise"] = ew
        line_extra["phases_ms"] = phases
        # ---- CPU baseline on the host cores (rank 0, N=1 only) ----
        if world == 1 and not args.no_cpu:
            psecs, pdesc = cpu_eval_seconds(inp, 15.0)
            got = ref_eval_seconds(inp, 15.0)
            secs, desc = got if got is not None else (psecs, pdesc)
            line_extra["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": host_threads(),
                                          "kind": "reference" if got is not None else "port",
                                          "sample": desc + ("; a full-size evaluation of the same library is what `bench.py --impl reference` times" if got is not None else ""),
                                          "host": host_env(),
                                          "oracle_port": {"value": 1.0 / psecs, "unit": UNIT, "sample": pdesc}}
    fit.close()

    # ---- config 3 (Student-t process, prior=TRUE): evaluation at the metric's size and the posterior-predictive
    # sampling of treatment effects (5 x 1000 multivariate-t draws over 747 units, R/kernel_TP_SE_class.R:112-124) ----
    tp_obj = None
    if rank == 0 and not args.no_fits:
        tinp = make_inputs(N_METRIC, P_METRIC, 0, kind="TP")
        tfit = _lib.Fit(lib, _lib.CS_TP, tinp["y"], tinp["X"], tinp["z"], tinp["w"], tinp["params"], inv_pipeline=args.inv_pipeline,
                        partition=args.partition)
        for _ in range(3):
            tfit.eval_async()
        tfit.sync()
        tfit.event_record(4)
        for _ in range(5):
            tfit.eval_async()
        tfit.event_record(5)
        tfit.sync()
        tp_ms = tfit.event_elapsed_ms(4, 5) / 5
        _, tst, _ = tfit.fetch()
        tfit.close()
        sinp = make_inputs(672, P_METRIC, 1, kind="TP")
        sfit = _lib.Fit(lib, _lib.CS_TP, sinp["y"], sinp["X"], sinp["z"], sinp["w"], sinp["params"])
        sfit.eval()
        tr = sfit.treat(sinp["X"], cov_jitter=1e-6, want_cov=True)
        lib.tp_sample(tr["cov"], 2000.0, 5000, 7, want_ate=True)
        t0 = time.perf_counter()
        U, aU = lib.tp_sample(tr["cov"], 2000.0, 5000, 7, want_ate=True)
        samp_ms = 1e3 * (time.perf_counter() - t0)
        sfit.close()
        tp_obj = {"what": "prior=TRUE: LML+gradient evaluation of the Student-t process at n=8192, p=25 (device-resident) and "
                          "cs_tp_sample on the 672 x 672 treatment covariance (5000 draws, 90 % order statistics, host buffers)",
                  "eval_ms": tp_ms, "evals_per_s": 1e3 / tp_ms, "lml": float(tst[1]), "tp_sample_ms": samp_ms,
                  "ate_U": float(aU), "U_mean": float(np.mean(U))}

    # ---- metric ii: replicate fits/hour (extra object) ----
    fits_obj = None
    if not args.no_fits:
        nfit = args.fits
        barrier()
        secs, nf, iters, pehe = measure_fits(lib, rank, world, nfit, args.fit_batch, args.host_sim)
        secs = max_over_ranks(secs)
        if dist is not None:
            all_it = [None] * world
            dist.all_gather_object(all_it, iters)
            iters = [i for sub in all_it for i in sub]
        fits_obj = {"value": world * nf / secs * 3600.0, "unit": "fits/hour", "fits": world * nf,
                    "iterations_mean": float(np.mean(iters)), "pehe_test_winners_rank0": pehe,
                    "what": "GP fit on a 672x25 Hill surface-B split to the reference stopping rule (maxiter 3000, tol 1e-1, "
                            "lr .01, Nadam) + predict(747) + treatment(672, 75); %d (replicate, restart) fits per GPU in batched handles of %d "
                            "members (cs_fit_create_batch, every launch serves a whole batch), handles driven concurrently "
                            "(cs_fit_run_many); %s; includes handle creation" % (nf, args.fit_batch, "host generator / metrics (--host-sim)" if args.host_sim
                                                                                else "simulation driver on the device (cs_sim_run)")}

    # ---- the two sharded configurations of BASELINE.json (extra objects; SCALE records them at N = 1, 2, 4, 8) ----
    c4_obj = c5_obj = None
    if not args.no_configs:
        c4_obj = measure_config4(lib, rank, world, dist, barrier, max_over_ranks, args.fit_batch, args.host_sim)
        if world in GRIDS:
            c5_obj = measure_config5(lib, torch, dev, rank, world, dist, barrier)

    stateless_obj = single_obj = None
    if rank == 0 and world == 1 and not args.no_fits:
        stateless_obj = measure_stateless_boundary()
        single_obj = measure_single_fit_configs()

    if rank == 0 and fits_obj and world == 1 and not args.no_cpu:
        # the same fit on the host: one reference-faithful optimizer iteration of the oracle port (para_update, which
        # builds Gram + inverse twice for the GP class, R/kernel_GP_SE_class.R:48,58) at the replicate size, timed
        # over 3 iterations and scaled by the mean iteration count measured above
        from oracle import causalstump_oracle as o
        sinp = make_inputs(672, P_METRIC, 1)
        Kc = o.KernelClass_GP_SE(P_METRIC, np.ones(672))
        Kc.parameters = o.unpack_params(sinp["params"], "GP", P_METRIC)
        opt = o.optNadam(0.01, 0.9, 0.999)
        opt.initOpt(Kc)
        Kc.para_update(1, sinp["y"], sinp["X"], sinp["z"], opt)
        t0 = time.perf_counter()
        for it in range(2, 5):
            Kc.para_update(it, sinp["y"], sinp["X"], sinp["z"], opt)
        sec_it = (time.perf_counter() - t0) / 3
        fits_obj["cpu_port"] = {"seconds_per_iteration": sec_it, "fits_per_hour_one_process": 3600.0 / (sec_it * fits_obj["iterations_mean"]),
                                "cores": host_threads(),
                                "what": "oracle port of para_update at n=672, p=25 (reference-faithful), scaled by iterations_mean"}
        from oracle import ref_binding as rb
        if rb.available():
            # the same iteration by the reference's own compiled routines, chained as para_update chains them:
            # kernmat + invkernel, grad, optimizer step, mu, then kernmat + invkernel + stats again (:48-58)
            ref = rb.RefLib()
            ref.use_host_lapack()
            par, ys, Xs, zs, ws = sinp["par"], sinp["y"], sinp["X"], sinp["z"], sinp["w"]
            flat = np.array(sinp["params"], dtype=np.float64)
            t0 = time.perf_counter()
            reps = 2
            for it in range(reps):
                for second in (False, True):
                    Kr, Kmr, Kar = ref.kernmat(0, Xs, Xs, zs, zs, par["Lm"], par["La"], par["lambdam"], par["lambdaa"])
                    invr = ref.invkernel(zs, ws, Kr, par["sigma"])
                    if not second:
                        gr, _ = ref.grad(0, ys, Xs, zs, ws, Kr, Kmr, Kar, invr, flat)
                        ref.opt_step(1, 0, P_METRIC, it + 1.0, 0.01, 0.9, 0.999, 1e-8, np.zeros_like(flat), np.zeros_like(flat), gr, flat)
                        ref.mu_solution(ys, zs, invr)
                    else:
                        ref.stats(0, P_METRIC, ys, Kr, invr, flat)
            sec_ref = (time.perf_counter() - t0) / reps
            fits_obj["cpu_reference"] = {"seconds_per_iteration": sec_ref,
                                         "fits_per_hour_one_process": 3600.0 / (sec_ref * fits_obj["iterations_mean"]),
                                         "cores": host_threads(),
                                         "what": "the reference's own C++ routines (oracle/_ref, host LAPACK) chained as one para_update "
                                                 "iteration at n=672, p=25, scaled by iterations_mean"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(), "parallelism": "replicas x%d (one per GPU), no collective on the data path" % world,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "lml": float(st[1]),
        }
        line.update(line_extra)
        check_fracs(line)
        if fits_obj:
            line["fits"] = fits_obj
        if tp_obj:
            line["student_t"] = tp_obj
        if stateless_obj:
            line["boundary_stateless"] = stateless_obj
        if single_obj:
            line["configs_1_2_3"] = single_obj
        if c4_obj:
            line["config4"] = c4_obj
        if c5_obj:
            line["config5"] = c5_obj
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fits", type=int, default=256, help="(replicate, restart) fits per GPU for the fits/hour object")
    ap.add_argument("--fit-batch", type=int, default=32, help="members per batched handle (cs_fit_create_batch); 1 = one handle per fit")
    ap.add_argument("--no-fits", action="store_true")
    ap.add_argument("--host-sim", action="store_true", help="replicate loop with the generator / split / metrics in host NumPy (round-1 driver) instead of cs_sim_run")
    ap.add_argument("--no-configs", action="store_true", help="skip the config4 (400 fits split over N) and config5 (n=32768 on a P x Q grid) objects")
    ap.add_argument("--inv-pipeline", type=int, default=1, help="1: inverse pipelined behind the Cholesky (default); 0: recursive inverse after it")
    ap.add_argument("--partition", type=int, default=0, help="SM partition for the Cholesky chain: 0 auto, 1 on, -1 off")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
