#!/usr/bin/env python
"""Per-launch timeline of the dense plan (Cholesky + pipelined inverse) of one n x p evaluation.
Prints per-stream busy time, the gaps on the critical stream S0 and the slowest launches.
    python tools/timeline.py [--n 8192] [--inv-pipeline 1] [--out gpurun_out/timeline.csv]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--p", type=int, default=25)
    ap.add_argument("--inv-pipeline", type=int, default=1)
    ap.add_argument("--partition", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--window", default="", help="t0,t1 in ms: list every launch that starts inside")
    a = ap.parse_args()
    import bench
    from causalstump_b200 import _lib
    lib = _lib.load()
    lib.require_device()
    inp = bench.make_inputs(a.n, a.p, 0)
    fit = _lib.Fit(lib, _lib.CS_GP, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], inv_pipeline=a.inv_pipeline, partition=a.partition)
    for _ in range(3):
        fit.eval_async()
    fit.sync()
    fit.timeline()
    tl = fit.timeline()
    if a.out:
        np.savetxt(a.out, tl, fmt="%.4f", delimiter=",", header="stream,type,tiles,K_or_blk,t0_ms,t1_ms")
    end = tl[:, 5].max()
    print("launches %d, span %.3f ms" % (len(tl), end))
    for s in range(8):
        m = tl[:, 0] == s
        if not m.any():
            continue
        d = tl[m, 5] - tl[m, 4]
        print("stream %d: %4d launches, busy %.3f ms (sum of event brackets), first %.3f last %.3f" %
              (s, m.sum(), d.sum(), tl[m, 4].min(), tl[m, 5].max()))
    m0 = tl[:, 0] == 0
    leaf = tl[(tl[:, 1] == 1)]
    dl = leaf[:, 5] - leaf[:, 4]
    print("leaf: %d launches, mean %.1f us, min %.1f, max %.1f" % (len(leaf), 1e3 * dl.mean(), 1e3 * dl.min(), 1e3 * dl.max()))
    # leaf-to-leaf period: the critical path per block step
    per = np.diff(leaf[:, 4])
    print("leaf-to-leaf period: mean %.1f us, median %.1f, max %.1f" % (1e3 * per.mean(), 1e3 * np.median(per), 1e3 * per.max()))
    W = 4
    inner = [per[i] for i in range(len(per)) if (i + 1) % W != 0]
    outer = [per[i] for i in range(len(per)) if (i + 1) % W == 0]
    print("  inside a panel: mean %.1f us;  across a panel boundary (incl. U1): mean %.1f us" %
          (1e3 * np.mean(inner), 1e3 * np.mean(outer)))
    if a.window:
        lo, hi = [float(x) for x in a.window.split(",")]
        for r in tl[(tl[:, 4] >= lo) & (tl[:, 4] < hi)]:
            print("S%d %-4s tiles %5d K/blk %5d  t0 %.4f t1 %.4f dur %7.1f us" %
                  (r[0], {0: "gemm", 1: "leaf", 4: "copy"}.get(int(r[1]), "?"), r[2], r[3], r[4], r[5], 1e3 * (r[5] - r[4])))
    s0 = tl[m0]
    print("S0 launches by duration (top 12):")
    d0 = s0[:, 5] - s0[:, 4]
    for i in np.argsort(-d0)[:12]:
        print("   type %d tiles %5d K/blk %5d  t0 %.3f dur %.1f us" % (s0[i, 1], s0[i, 2], s0[i, 3], s0[i, 4], 1e3 * d0[i]))
    fit.close()


if __name__ == "__main__":
    main()
