"""Regenerates tests/golden/golden_ref_8192.npz: ONE evaluation of the reference's OWN C++ at the headline size.

Every number in the fixture is produced by oracle/_ref/libcausalstump_ref.so (the reference's
src/kernel_GP_SE_cpp.cpp / src/kernel_TP_SE_cpp.cpp compiled unmodified, oracle/Makefile), chained the way
para_update chains the routines (R/kernel_GP_SE_class.R:45-60): kernmat -> invkernel -> grad -> mu_solution.
arma::inv_sympd / arma::log_det go through the host LAPACK (dpotrf + dpotri, dgetrf: what Armadillo links).

Cases (inputs = bench.make_inputs, the generator of the timed workload, so the test re-creates them from the seed;
the fixture stores a SHA-256 of the input bytes and the test refuses to compare if the generator drifted):
  gp8192  GP (prior=FALSE), n=8192, p=25, replicate 0            -- the configuration `value` is quoted on
  tp8192  Student-t process (prior=TRUE, nu=2000), n=8192, p=25  -- config 3 at the metric's size
  gp4000  GP, n=4000 (ragged: padded to 4096, the smallest size with the SM partition / depth-3 plan), weights

Stored per case: gradient (2p+4 / 2p+5), stats (squared residual, LML), mu_solution, diag(K^-1), rows
{0, n/4+11, n/2+1, n-1} of K^-1, the parameter vector and the input digest.  ~0.5 MB in total.

    python tests/golden/make_golden_ref_8192.py     # build container only (needs /root/reference); ~10 minutes
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ref_8192.npz")
CASES = [("gp8192", "GP", 8192, False), ("tp8192", "TP", 8192, False), ("gp4000", "GP", 4000, True)]


def case_inputs(kind, n, weights):
    inp = bench.make_inputs(n, bench.P_METRIC, 0, kind=kind)
    w = np.random.default_rng(4000).uniform(0.5, 2.0, n) if weights else np.ones(n)
    return inp["y"], inp["X"], inp["z"], w, np.asarray(inp["params"], dtype=np.float64)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
    return h.hexdigest()


def rows_of(n):
    return np.array([0, n // 4 + 11, n // 2 + 1, n - 1])


def build(only=None):
    if rb.build() is None:
        raise SystemExit("oracle/_ref cannot be built here (no /root/reference)")
    ref = rb.RefLib()
    lap = ref.use_host_lapack()
    out = {}
    if os.path.exists(OUT):
        out.update(dict(np.load(OUT)))
    for name, kind, n, weights in CASES:
        if only and name not in only:
            continue
        y, X, z, w, par = case_inputs(kind, n, weights)
        t0 = time.perf_counter()
        g, st, mu, mats = ref.eval_lml_grad(0 if kind == "GP" else 1, y, X, z, w, par)
        secs = time.perf_counter() - t0
        rows = rows_of(n)
        pre = name + "/"
        out.update({
            pre + "meta": np.array([0 if kind == "GP" else 1, n, X.shape[1], int(weights)]),
            pre + "params": par, pre + "grad": g, pre + "stats": st, pre + "mu_sol": np.array(mu),
            pre + "invK_diag": np.diag(mats["invK"]).copy(), pre + "rows": rows, pre + "invK_rows": mats["invK"][rows].copy(),
            pre + "digest": np.array(digest(y, X, z, w, par)), pre + "ref_seconds": np.array(secs),
            pre + "lapack": np.array(lap),
        })
        print("%s: LML %.12e  |grad|max %.4e  mu %.12e  %.1f s" % (name, st[1], np.max(np.abs(g)), mu, secs), flush=True)
        del mats
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    build(sys.argv[1:] or None)
