#!/usr/bin/env python
"""Per-launch timeline of the dense plan of ONE batched handle (B fits of n x p): where a batched small-n iteration's
Cholesky / inverse time goes, launch by launch.    python tools/batch_timeline.py [B] [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import bench
    from causalstump_b200 import _lib
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 672
    lib = _lib.load()
    lib.require_device()
    members = []
    for b in range(B):
        inp = bench.make_inputs(n, 25, b)
        members.append((inp["y"], inp["X"], inp["z"], None, inp["params"]))
    fb = _lib.FitBatch(lib, 0, members)
    for _ in range(3):
        fb.eval_async()
    fb.sync()
    fb.timeline()
    tl = fb.timeline()
    print("B=%d n=%d: %d launches, span %.3f ms" % (B, n, len(tl), tl[:, 5].max()))
    names = {0: "gemm", 1: "leaf", 4: "copy"}
    agg = {}
    for r in tl:
        key = (names.get(int(r[1]), "?"), int(r[3]) if int(r[1]) == 0 else 0)
        a = agg.setdefault(key, [0, 0.0, 0])
        a[0] += 1; a[1] += r[5] - r[4]; a[2] += int(r[2])
    for (nm, K), (cnt, ms, tiles) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-5s K=%4d  launches %3d  tiles/member %5d  total %.3f ms  avg %.1f us" % (nm, K, cnt, tiles, ms, 1e3 * ms / cnt))
    gaps = tl[1:, 4] - tl[:-1, 5]
    print("launch gaps (same stream, consecutive): total %.3f ms, mean %.1f us" % (gaps[gaps > 0].sum(), 1e3 * gaps[gaps > 0].mean()))
    fb.close()


if __name__ == "__main__":
    main()
