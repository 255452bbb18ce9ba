"""CPU, world_size 2 over gloo: the N > 1 plumbing (single broadcast of the sources, disjoint exhaustive sharding)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pixelsynth_b200.parallel import broadcast_sources, view_assignment

    src = torch.arange(2 * 3 * 8 * 8, dtype=torch.float32).view(2, 3, 8, 8) if rank == 0 else torch.zeros(2, 3, 8, 8)
    broadcast_sources(src, world)
    pairs = view_assignment(4, 8, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (float(src.sum()), pairs))
    if rank == 0:
        ret["ok"] = (all(abs(g[0] - gathered[0][0]) < 1e-6 for g in gathered) and gathered[0][0] > 0,
                     sorted(p for g in gathered for p in g[1]))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    same, pairs = ret["ok"]
    assert same
    assert pairs == sorted((i, v) for i in range(4) for v in range(8))     # every pair exactly once


def test_shard_ranges():
    from pixelsynth_b200.parallel import shard_range, view_assignment

    for n, w in ((64, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert len(view_assignment(256, 1, 3, 8)) == 32        # config 5: by image
    assert view_assignment(64, 8, 5, 8) == [(i, 5) for i in range(64)]   # config 4: GPU g renders view g
