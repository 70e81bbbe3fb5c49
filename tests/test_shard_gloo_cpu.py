"""World-size-2 (and 3) gloo tests of the N>1 path's host logic: views are sharded round-robin,
LOS rays contiguously, and only compact results are gathered.  Runs on CPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, n_rays, q):
    for p in (ROOT, os.path.join(ROOT, "pixelmek-3d_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from pm3d import shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # per-view compact result (e.g. a frame checksum) = f(view index); every rank computes its shard
        mine = shard.view_shard(n_views, rank, world)
        local = torch.from_numpy((mine * 7919 + 13).astype(np.int64))
        full = shard.gather_views(local, n_views, rank, world)
        want = torch.arange(n_views, dtype=torch.int64) * 7919 + 13
        ok_views = bool(torch.equal(full, want))
        # per-ray result: 1 byte per ray, contiguous shards of uneven size
        lo, hi = shard.ray_shard(n_rays, rank, world)
        local_r = torch.from_numpy(((np.arange(lo, hi) * 31) % 2).astype(np.uint8))
        full_r = shard.gather_rays(local_r, n_rays, rank, world)
        want_r = torch.from_numpy(((np.arange(n_rays) * 31) % 2).astype(np.uint8))
        ok_rays = bool(torch.equal(full_r, want_r))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok_views, ok_rays, float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_views,n_rays", [(2, 48, 1001), (2, 7, 5), (3, 10, 100)])
def test_sharded_gather(world, n_views, n_rays):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, n_rays, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_views, ok_rays, tmax in results:
        assert ok_views and ok_rays, f"rank {rank}"
        assert tmax == float(world)
