"""world_size-2 gloo test of the replicate/restart sharding (the N>1 path has no data-path
collective; only the result gather uses torch.distributed)."""
import os
import socket

import pytest

from causalstump_b200 import replicates as rp


def fake_unit(rep, rs):
    return dict(replicate=rep, restart=rs, lml=-100.0 + ((rep * 7 + rs * 3) % 5), iters=10 + rep + rs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allr, local = rp.run_sharded(fake_unit, 5, 3, rank, world, dist)
    q.put((rank, allr, local))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_every_unit_once():
    units = rp.unit_list(100, 4)
    for world in (1, 2, 4, 8):
        seen = sorted(u for r in range(world) for u in rp.shard(units, r, world))
        assert seen == sorted(units)
        sizes = [len(rp.shard(units, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_winner_selection():
    res = [fake_unit(rep, rs) for rep, rs in rp.unit_list(4, 3)]
    win = rp.select_winners(res)
    assert [w["replicate"] for w in win] == [0, 1, 2, 3]
    for w in win:
        assert w["lml"] == max(r["lml"] for r in res if r["replicate"] == w["replicate"])


def test_gloo_world2_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        rank, allr, local = q.get(timeout=120)
        out[rank] = (allr, local)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single, _ = rp.run_sharded(fake_unit, 5, 3)
    assert out[0][0] == single and out[1][0] is None
    assert len(out[0][1]) + len(out[1][1]) == 15
    assert rp.select_winners(out[0][0]) == rp.select_winners(single)
