#!/usr/bin/env python
"""Throughput of n=8192 evaluations with 1, 2 or 3 independent handles in flight on one GPU (replicas of the metric's workload
sharing the device: the idle phases of one evaluation -- start-up of the first panel, the chain partition, the gradient tail --
are filled by another).    python tools/two_in_flight.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from causalstump_b200 import _lib
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    lib = _lib.load()
    lib.require_device()
    for nh in (1, 2, 3):
        fits = []
        for h in range(nh):
            inp = bench.make_inputs(8192, 25, h)
            fits.append(_lib.Fit(lib, 0, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], inv_pipeline=1))
        for _ in range(3):
            for f in fits:
                f.eval_async()
        for f in fits:
            f.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            for f in fits:
                f.eval_async()
        for f in fits:
            f.sync()
        dt = time.perf_counter() - t0
        lml = [f.fetch()[1][1] for f in fits]
        print("%d handle(s) in flight: %d evaluations in %.3f s = %.2f evals/s (%.2f ms per evaluation), LML %s" %
              (nh, steps * nh, dt, steps * nh / dt, 1e3 * dt / (steps * nh), ["%.6f" % v for v in lml]), flush=True)
        for f in fits:
            f.close()


if __name__ == "__main__":
    main()
