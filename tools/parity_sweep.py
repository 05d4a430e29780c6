#!/usr/bin/env python
"""Randomised parity sweep on the GPU: evaluation (gradient, LML, mu, squared residual) of random (kind, n, p, weights, tile,
plan) cases against the oracle, plus short optimisation runs; prints the worst relative errors against the tolerances of
BASELINE.json (LML 1e-9, gradients 1e-8)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from causalstump_b200 import _lib
    from oracle import causalstump_oracle as o
    from tests import _cases as cs
    lib = _lib.load()
    lib.require_device()
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    use_ref = "--ref" in sys.argv   # check against the reference's own C++ (oracle/_ref) instead of the NumPy oracle
    ref = None
    if use_ref:
        from oracle import ref_binding as rb
        assert rb.available(), "oracle/_ref/libcausalstump_ref.so not present"
        ref = rb.RefLib()
    rng = np.random.default_rng(int(args[0]) if len(args) > 0 else 20260101)
    ncase = int(args[1]) if len(args) > 1 else 48
    worst = dict(grad=0.0, lml=0.0, mu=0.0, res=0.0, fit_par=0.0, fit_lml=0.0)
    t0 = time.time()
    for c in range(ncase):
        kind = "GP" if rng.random() < 0.6 else "TP"
        sizes = [2, 3, 17, 63, 64, 65, 127, 128, 129, 200, 333, 512, 640, 747] + ([] if use_ref else [1000, 1300, 1537])
        n = int(rng.choice(sizes))
        p = int(rng.choice([1, 2, 5, 13, 25, 26, 32, 33, 40, 60, 100]))
        tile = int(rng.choice([0, 0, 64, 128]))
        pipe = int(rng.integers(0, 2))
        weights = bool(rng.random() < 0.5)
        P = cs.make_problem(kind, n, p, replicate=int(rng.integers(0, 1000)), weights=weights, nu=float(rng.choice([2.0, 50.0, 2000.0])),
                            draws=(float(rng.uniform(0.2, 1.0)), float(rng.uniform(0.2, 1.0))))
        if use_ref:
            gflat, st, mu, _ = ref.eval_lml_grad(cs.kcode(kind), P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]))
            g = o.unpack_params(gflat, kind, P["p"])
        else:
            g, st, mu = o.eval_lml_grad(kind, P["y"], P["X"], P["z"], P["w"], P["par"])
        f = _lib.Fit(lib, cs.kcode(kind), P["y"], P["X"], P["z"], P["w"], o.pack_params(P["par"]), tile=tile, inv_pipeline=pipe)
        g2, st2, mu2 = f.eval()
        gref = o.pack_params(g)
        e_g = float(np.max(np.abs(g2 - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref)))))
        e_l = abs(st2[1] - st[1]) / abs(st[1])
        e_m = abs(mu2 - mu) / max(abs(mu), 1e-300)
        e_r = abs(st2[0] - st[0]) / max(abs(st[0]), 1e-300)
        worst["grad"] = max(worst["grad"], e_g); worst["lml"] = max(worst["lml"], e_l)
        worst["mu"] = max(worst["mu"], e_m); worst["res"] = max(worst["res"], e_r)
        line = "case %2d %s n=%4d p=%2d tile=%3d pipe=%d w=%d  grad %.1e lml %.1e mu %.1e res %.1e" % (c, kind, n, p, tile, pipe, weights, e_g, e_l, e_m, e_r)
        if n <= 400 and c % 3 == 0:  # a short optimisation run, same number of steps on both sides
            fit = cs.oracle_fit(P, "Nadam", 6, 1e-14, 0.01)
            f.set_params(o.pack_params(P["par"]))
            f.reset_optimizer()
            it, tr = f.run(6, 1e-14)
            e_p = cs.rel(f.get_params(), o.pack_params(fit.Kernel.parameters), 1e-9)
            e_t = cs.rel(tr[1, 1:], fit.stats[1, 1:])
            worst["fit_par"] = max(worst["fit_par"], e_p); worst["fit_lml"] = max(worst["fit_lml"], e_t)
            line += "  | 6 steps: params %.1e lml-trace %.1e" % (e_p, e_t)
        f.close()
        print(line, flush=True)
        assert e_g < 1e-8 and e_l < 1e-9, "tolerance exceeded"
    print("worst over %d cases: %s  (%.0f s)" % (ncase, {k: "%.2e" % v for k, v in worst.items()}, time.time() - t0))


if __name__ == "__main__":
    main()
