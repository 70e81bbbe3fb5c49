"""Multi-GPU side of the C-ABI (csrc/comm.cu): sharded view / LOS batches whose compact results are gathered by
a collective the library enqueues itself.  One-GPU boxes run the single-rank communicator and a one-device group;
the two-rank and two-device cases need `gpurun --gpus 2` and are skipped otherwise."""
import threading

import numpy as np
import pytest

import pm3d
from pm3d import shard
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


def n_devices():
    return pm3d.load().pm3d_device_count()


def small_batch(n_views=37, w=160, h=100):
    cfg = W.config_c4(n_views, w, h, k_sprites=20)
    return cfg, cfg["views"], cfg["sprites"], cfg["offsets"]


def single_gpu_results(ctx, cfg, views, sprites, offsets, w, h):
    ctx.upload_scene(cfg["scene"])
    ctx.set_view(w, h, len(views), len(sprites))
    res = ctx.render(views, sprites, offsets, want_frame=False)
    return res["conv"].copy(), ctx.frame_checksums(len(views)).copy()


def los_case(n=100003, seed=5):
    grid = W.synthetic_levels(96, 96, 1, 0.08, 9)
    segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
    return segs, W.los_rays(grid[0], n, seed)


def run_rank(rank, world, uid, cfg, views, sprites, offsets, w, h, segs, rays, out):
    """What one process-per-GPU rank does, through device pointers only."""
    try:
        ctx = pm3d.Context(rank)
        ctx.comm_init(uid, rank, world)
        assert (ctx.comm_rank, ctx.comm_size) == (rank, world)
        mine = shard.view_shard(len(views), rank, world)
        assert len(mine) == pm3d.load().pm3d_shard_view_count(len(views), rank, world)
        lv = views[mine]
        recs = [sprites[offsets[i]:offsets[i + 1]] for i in mine]
        lo_ = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.int32)
        ls = np.concatenate(recs) if recs else sprites[:0]
        ctx.upload_scene(cfg["scene"])
        ctx.set_view(w, h, max(len(lv), 1), max(len(ls), 1))
        nv = len(views)
        d = dict(v=ctx.device_alloc(max(lv.nbytes, 8)), s=ctx.device_alloc(max(ls.nbytes, 8)), o=ctx.device_alloc(lo_.nbytes),
                 out=ctx.device_alloc(max(len(lv), 1) * w * h * 4), conv=ctx.device_alloc(nv * 32), sums=ctx.device_alloc(nv * 8))
        if len(lv):
            ctx.h2d(d["v"], lv); ctx.h2d(d["o"], lo_)
        if len(ls):
            ctx.h2d(d["s"], ls)
        for _ in range(3):   # repeated: the gather staging is reused
            ctx.comm_render_views(d["v"], len(lv), d["s"], d["o"], len(ls), d["out"], nv, d["conv"], d["sums"])
        ctx.sync()
        conv = np.zeros(nv, dtype=pm3d._lib.CONV_DTYPE)
        sums = np.zeros(nv, dtype=np.uint64)
        ctx.d2h(conv, d["conv"]); ctx.d2h(sums, d["sums"])
        # LOS: contiguous range, results of all ranks everywhere
        ctx.upload_segments(segs)
        lo, hi = shard.ray_shard(len(rays), rank, world)
        d_r, d_all = ctx.device_alloc(max((hi - lo) * 32, 8)), ctx.device_alloc(len(rays))
        if hi > lo:
            ctx.h2d(d_r, rays[lo:hi])
        ctx.comm_los_batch(d_r, hi - lo, len(rays), d_all)
        ctx.sync()
        vis = np.zeros(len(rays), dtype=np.uint8)
        ctx.d2h(vis, d_all)
        # a batch whose shards are whole 32-ray words travels as bits (k_pack_bits / k_unpack_bits); twice: staging reuse
        n2 = 98304
        lo, hi = shard.ray_shard(n2, rank, world)
        ctx.h2d(d_r, rays[lo:hi])
        for _ in range(2):
            ctx.comm_los_batch(d_r, hi - lo, n2, d_all)
        ctx.sync()
        vis2 = np.zeros(n2, dtype=np.uint8)
        ctx.d2h(vis2, d_all)
        out[rank] = (conv, sums, vis, vis2)
        ctx.close()
    except Exception as e:  # surfaced by the caller
        out[rank] = e


@pytest.mark.parametrize("world", [1, 2, 3])
def test_comm_ranks_gather_every_view_and_ray(gpu_ctx, world, monkeypatch):
    if n_devices() < world:
        pytest.skip(f"needs {world} GPUs")
    if world == 1:
        monkeypatch.setenv("PM3D_COMM_FORCE_BITS", "1")   # one rank packs / unpacks its own answers too
    w, h = 160, 100
    cfg, views, sprites, offsets = small_batch(37, w, h)
    want_conv, want_sums = single_gpu_results(gpu_ctx, cfg, views, sprites, offsets, w, h)
    segs, rays = los_case()
    gpu_ctx.upload_segments(segs)
    want_vis = gpu_ctx.los_batch(rays)
    uid = pm3d.comm_unique_id()
    out = [None] * world
    ts = [threading.Thread(target=run_rank, args=(r, world, uid, cfg, views, sprites, offsets, w, h, segs, rays, out)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    for r in range(world):
        if isinstance(out[r], Exception):
            raise out[r]
        conv, sums, vis, vis2 = out[r]
        assert conv.tobytes() == want_conv.tobytes(), f"rank {r}: convergence records differ"
        assert np.array_equal(sums, want_sums), f"rank {r}: checksums differ"
        assert np.array_equal(vis, want_vis), f"rank {r}: LOS answers differ"
        assert np.array_equal(vis2, want_vis[:len(vis2)]), f"rank {r}: LOS answers of the bit-packed gather differ"
    assert 0 < want_vis.sum() < len(rays)


@pytest.mark.parametrize("n_dev", [1, 2, 4])
def test_group_one_process_many_gpus(gpu_ctx, n_dev):
    if n_devices() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    w, h = 160, 100
    cfg, views, sprites, offsets = small_batch(37, w, h)
    want_conv, want_sums = single_gpu_results(gpu_ctx, cfg, views, sprites, offsets, w, h)
    segs, rays = los_case(50001)
    gpu_ctx.upload_segments(segs)
    want_vis = gpu_ctx.los_batch(rays)
    g = pm3d.Group(list(range(n_dev)))
    try:
        def setup(c):
            c.upload_scene(cfg["scene"])
            c.set_view(w, h, (len(views) + n_dev - 1) // n_dev, len(sprites))
            c.upload_segments(segs)
        g.each(setup)
        for _ in range(2):
            conv, sums = g.render_views(views, sprites, offsets)
            assert conv.tobytes() == want_conv.tobytes() and np.array_equal(sums, want_sums)
            assert np.array_equal(g.los_batch(rays), want_vis)
        # fewer views than devices, no sprites
        c1, s1 = g.render_views(views[:1], None, None)
        gpu_ctx.render(views[:1], None, want_frame=False)
        assert s1[0] == gpu_ctx.frame_checksums(1)[0]
    finally:
        g.close()


def test_comm_calls_need_a_communicator(gpu_ctx):
    with pytest.raises(pm3d.Pm3dError, match="pm3d_comm_init"):
        gpu_ctx.comm_los_batch(0, 0, 16, 0)
    assert gpu_ctx.comm_size == 1 and gpu_ctx.comm_rank == 0
