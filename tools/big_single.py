#!/usr/bin/env python
"""n = 32768 (or argv[1]) evaluation on the single-GPU engine: ms per evaluation for the current CS_* environment."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from causalstump_b200 import _lib
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    lib = _lib.load()
    inp = bench.make_inputs(n, 25, 0)
    f = _lib.Fit(lib, 0, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], inv_pipeline=1)
    f.eval_async(); f.sync()
    f.event_record(6)
    for _ in range(2):
        f.eval_async()
    f.event_record(7)
    f.sync()
    ms = f.event_elapsed_ms(6, 7) / 2
    print("n=%d CS_TAIL_A=%s CS_PANEL=%s: %.1f ms per evaluation = %.2f TFLOP/s (n^3), LML %.6f" %
          (n, os.environ.get("CS_TAIL_A"), os.environ.get("CS_PANEL"), ms, float(n) ** 3 / (ms * 1e-3) / 1e12, f.fetch()[1][1]), flush=True)
    f.close()


if __name__ == "__main__":
    main()
