#!/usr/bin/env python
"""Per-phase device time of one batched evaluation (B fits of n x p) -- where a small-n batch iteration goes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import bench
    from causalstump_b200 import _lib
    lib = _lib.load()
    lib.require_device()
    n, p = 672, 25
    for B in (1, 8, 32, 64):
        for tile in (64, 128):
            members = []
            for b in range(B):
                inp = bench.make_inputs(n, p, b)
                members.append((inp["y"], inp["X"], inp["z"], None, inp["params"]))
            fb = _lib.FitBatch(lib, 0, members, tile=tile)
            for _ in range(3):
                fb.eval_async()
            fb.sync()
            ph = None
            for _ in range(5):
                q = fb.eval_profile()
                ph = q if ph is None else {k: min(ph[k], q[k]) for k in q}
            tot = sum(ph.values())
            print("B=%2d tile=%3d total %.3f ms (%.1f us/fit)  " % (B, tile, tot, 1e3 * tot / B) +
                  " ".join("%s=%.3f" % (k, v) for k, v in ph.items()), flush=True)
            fb.close()


if __name__ == "__main__":
    main()
