"""World-size-2 gloo test (CPU) of the scene sharding and the final result gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmwave_msc_b200 import sharding


def test_shard_bounds_cover_everything():
    for n, w in ((8192, 8), (1000, 3), (5, 8), (0, 2)):
        ids = []
        for r in range(w):
            lo, hi = sharding.shard_bounds(n, w, r)
            assert 0 <= hi - lo <= -(-n // w) if n else hi == lo
            ids += list(range(lo, hi))
        assert ids == list(range(n))


def _worker(rank, world, port, n_scenes, per_scene, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.shard_scene_ids(n_scenes, world, rank)
    # each scene's packed result row is a function of its global id only
    local = torch.tensor([[s * 1000 + j for j in range(per_scene)] for s in ids], dtype=torch.float32).reshape(-1)
    out = sharding.gather_results(local, n_scenes, per_scene)
    exp = torch.tensor([[s * 1000 + j for j in range(per_scene)] for s in range(n_scenes)],
                       dtype=torch.float32).reshape(-1)
    q.put((rank, bool(torch.equal(out, exp))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
