#!/usr/bin/env python
"""Config 5 over real ranks (NCCL): parity at a size the oracle handles, then timing of the
large single fit (default n=32768, p=25).  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/dist_check.py [--size 32768] [--steps 2] [--paritysize 1500]

Rank 0 prints one JSON line.  Grid: N=1 1x1, 2 2x1, 4 2x2, 8 2x4."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GRIDS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--dims", type=int, default=25)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--paritysize", type=int, default=1500)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from causalstump_b200 import _lib
    import bench
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lib.check(lib.dll.cs_set_device(local))
    P, Q = GRIDS[world]

    def make(n):
        inp = bench.make_inputs(n, args.dims, 0)
        ids = [lib.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return inp, _lib.DistFit(lib, 0, inp["y"], inp["X"], inp["z"], inp["w"], inp["params"], rank, world, P, Q, ids[0])

    out = {"what": "config 5: distributed LML+gradient evaluation, 2-D block-cyclic over a %dx%d grid, NCCL" % (P, Q),
           "n_gpus": world}
    # ---- parity against the oracle (rank 0 computes the oracle) ----
    if args.paritysize > 0:
        inp, f = make(args.paritysize)
        g, st, mu, _ = f.eval()
        f.close()
        if rank == 0:
            from oracle import causalstump_oracle as o
            go, sto, muo = o.eval_lml_grad("GP", inp["y"], inp["X"], inp["z"], inp["w"], inp["par"])
            gref = o.pack_params(go)
            out["parity"] = {"n": args.paritysize,
                             "grad_rel": float(np.max(np.abs(g - gref) / np.maximum(np.abs(gref), 1e-6 * np.max(np.abs(gref))))),
                             "lml_rel": float(abs(st[1] - sto[1]) / abs(sto[1])), "mu_rel": float(abs(mu - muo) / abs(muo))}
        gs = [None] * world
        dist.all_gather_object(gs, (g.tolist(), st.tolist()))
        if rank == 0:
            out["parity"]["ranks_agree"] = all(np.allclose(gs[0][0], x[0], rtol=1e-12, atol=0) for x in gs)
    # ---- timing of the large single fit ----
    inp, f = make(args.size)
    f.eval()  # warm-up
    dist.barrier(); torch.cuda.synchronize(dev)
    ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g, st, mu, m = f.eval()
        ms.append(m)
    torch.cuda.synchronize(dev); dist.barrier()
    wall = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([float(np.mean(ms)), wall * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    f.close()
    if rank == 0:
        n = args.size
        out.update({"n": n, "p": args.dims, "ms_per_eval_device_max": float(t[0]), "ms_per_eval_wall_max": float(t[1]),
                    "evals_per_s": 1e3 / float(t[1]), "tflops_n3": float(n) ** 3 / (float(t[1]) * 1e-3) / 1e12,
                    "lml": float(st[1]), "finite": bool(np.isfinite(g).all() and np.isfinite(st).all())})
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
