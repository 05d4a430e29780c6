#!/usr/bin/env python
"""Metric ii: fits/hour as a function of the number of concurrent fits on one GPU."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    from causalstump_b200 import _lib, replicates as rp
    lib = _lib.load()
    lib.require_device()
    rp.hill_units_concurrent([(0, 0)])
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    batch, tile, chunk = 32, 0, 0
    device = "--device" in sys.argv
    for a in sys.argv[1:]:
        if a.startswith("--batch="):
            batch = int(a.split("=")[1])
        if a.startswith("--tile="):
            tile = int(a.split("=")[1])
        if a.startswith("--chunk="):
            chunk = int(a.split("=")[1])
    rp.hill_units_concurrent([(0, 0), (0, 1)], batch=batch, tile=tile)
    for nf in [int(a) for a in args] or [4, 16, 32, 64]:
        units = [(k // 4, k % 4) for k in range(nf)]
        l0 = lib.launch_count()
        t0 = time.perf_counter()
        res = rp.hill_units_device(units, batch=batch) if device else rp.hill_units_concurrent(units, batch=batch, tile=tile, chunk=chunk)
        s = time.perf_counter() - t0
        it = [r["iters"] for r in res]
        print(json.dumps({"fits": nf, "driver": "device (cs_sim_run)" if device else "host generator / metrics", "batch": batch, "tile": tile, "chunk": chunk, "seconds": s, "fits_per_hour": nf / s * 3600, "iters_mean": float(np.mean(it)),
                          "iters_total": int(np.sum(it)), "us_per_fit_iter": 1e6 * s / np.sum(it), "launches": lib.launch_count() - l0,
                          "phases_s": {k: round(v, 3) for k, v in rp.LAST_TIMING.items()}}), flush=True)


if __name__ == "__main__":
    main()
