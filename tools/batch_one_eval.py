#!/usr/bin/env python
"""One batched evaluation (B fits of 672 x 25) after a warm-up one -- the target of the small-n ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from causalstump_b200 import _lib
    lib = _lib.load()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    members = []
    for b in range(B):
        inp = bench.make_inputs(672, 25, b)
        members.append((inp["y"], inp["X"], inp["z"], None, inp["params"]))
    fb = _lib.FitBatch(lib, 0, members)
    for _ in range(2):
        fb.eval_async()
    fb.sync()
    print("done", fb.select(0).fetch()[1])
    fb.close()


if __name__ == "__main__":
    main()
