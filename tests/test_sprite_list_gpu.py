"""GPU parity tests of pm3d_raycast_sprites_batch (csrc/sprite_list_kernel.cu) against the oracle's restatement of
Game.getRaycastSprites, and of the device-resident chain camera poses -> clutter -> sprite lists -> frames."""
import numpy as np
import pytest

import pm3d
from conftest import orc_camera
from oracle import oracle as O
from pm3d import clutter as cl
from pm3d import sprites as sp
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rd,n_world", [(25.0, 400), (2.3, 400), (-1.0, 77), (100.0, 1), (7.5, 0), (0.0, 130)])
def test_lists_equal_the_oracle(gpu_ctx, rd, n_world):
    scene, g = W.proving_grounds_scene()
    world = W.golden_sprites(scene, g)[:n_world]
    rng = W.SplitMix64(int(rd * 10) + n_world + 1000)
    n = 70
    u = rng.uniform(n * 4).reshape(n, 4)
    cams = [(500 * a - 0.5, 500 * b - 0.5) for a, b, _, _ in u]
    if n_world:
        cams[3] = (float(world[0]["x"]) - np.ceil(rd), float(world[0]["y"]))        # exactly on the proximity border
        cams[4] = (float(world[0]["x"]) - np.ceil(rd) - 1e-9, float(world[0]["y"]))  # just outside
    targets = [int(c * (n_world + 1)) - 1 for _, _, c, _ in u]
    clutter_n = np.array([int(d * 9) for _, _, _, d in u], dtype=np.int32)
    clutter_n[5] = 12                                                                # more than the stride: clamped
    clutter = np.resize(W.golden_sprites(scene, g)[::-1], n * 8).reshape(n, 8).copy()
    extras = [W.golden_sprites(scene, g)[i:i + (i % 3)] for i in range(n)]
    q = sp.list_queries(cams, targets)
    got, off, idx = gpu_ctx.raycast_sprites_batch(world, q, rd, (clutter_n, clutter), extras)
    assert off[0] == 0 and off[-1] == len(got)
    for i in range(n):
        want, widx = O.raycast_sprites(world.astype(O.SPRITE_DTYPE), cams[i][0], cams[i][1], rd, targets[i],
                                       clutter[i, :min(clutter_n[i], 8)].astype(O.SPRITE_DTYPE), extras[i].astype(O.SPRITE_DTYPE))
        assert off[i + 1] - off[i] == len(want), f"camera {i}"
        assert got[off[i]:off[i + 1]].tobytes() == want.tobytes(), f"camera {i}"
        assert np.array_equal(idx[off[i]:off[i + 1]], widx)
    # without clutter / extras
    got2, off2, idx2 = gpu_ctx.raycast_sprites_batch(world, q, rd)
    for i in range(n):
        want, widx = O.raycast_sprites(world.astype(O.SPRITE_DTYPE), cams[i][0], cams[i][1], rd, targets[i])
        assert got2[off2[i]:off2[i + 1]].tobytes() == want.tobytes() and np.array_equal(idx2[off2[i]:off2[i + 1]], widx)


def test_capacity_and_arguments(gpu_ctx):
    scene, g = W.proving_grounds_scene()
    world = W.golden_sprites(scene, g)
    q = sp.list_queries([(30.0, 30.0), (31.0, 29.0)])
    got, off, _ = gpu_ctx.raycast_sprites_batch(world, q, -1.0)
    assert off.tolist() == [0, len(world), 2 * len(world)]
    with pytest.raises(pm3d.Pm3dError, match="max_total"):
        gpu_ctx.raycast_sprites_batch(world, q, -1.0, max_total=2 * len(world) - 1)
    with pytest.raises(pm3d.Pm3dError):
        gpu_ctx.raycast_sprites_batch(world, q, float("nan"))
    one, idx = sp.getRaycastSprites(gpu_ctx, world, 30.0, 30.0, 10.0, target=len(world) - 1, clutter=world[:3], extras=world[5:6])
    want, widx = O.raycast_sprites(world.astype(O.SPRITE_DTYPE), 30.0, 30.0, 10.0, len(world) - 1, world[:3].astype(O.SPRITE_DTYPE),
                                   world[5:6].astype(O.SPRITE_DTYPE))
    assert one.tobytes() == want.tobytes() and np.array_equal(idx, widx)


def test_resident_chain_from_poses_to_frames(gpu_ctx):
    """Camera poses -> pm3d_clutter_batch -> pm3d_raycast_sprites_batch -> pm3d_render_resident with every
    intermediate in HBM; the frames equal the oracle's frames built from the oracle's clutter and sprite lists."""
    ctx = gpu_ctx
    L = pm3d._lib
    scene, g = W.proving_grounds_scene()
    scene.sheets = list(scene.sheets) + [W.sprite_sheet(18, 18, 41), W.sprite_sheet(24, 24, 42)]
    i_rock, i_bush = len(scene.sheets) - 2, len(scene.sheets) - 1
    ctx.upload_scene(scene)
    clutter = [cl.MapClutter(str(im), float(f), float(h), str(m) or None) for im, m, f, h in
               zip(g["clutter_image"], g["clutter_match"], g["clutter_frequency"], g["clutter_height"])]
    sheets = {clutter[0].image: (i_rock, 18, 18), clutter[1].image: (i_bush, 24, 24)}
    defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(s) for s in g["floor_images"]], sheets)
    ctx.upload_clutter(defs, cls, masks, int(g["map_seed"]))
    world = W.golden_sprites(scene, g)
    n, w, h, rd, cdist = 6, 256, 144, 40.0, 12.0
    poses = [(31.5, 3.5, 0.5, W.cardinal_to_angle(45.0), 0.0)] + W.free_cell_poses(scene, n - 1, 99)
    poses = [(p[0], p[1], 0.5, p[3], 0.0) for p in poses]
    views = W.make_views(scene, w, h, poses, render_distance=rd)
    cq = cl.clutter_queries([(p[0], p[1], p[3]) for p in poses], 70.0)
    lq = sp.list_queries([(p[0], p[1]) for p in poses])
    stride = 21 * (int(cdist) + 1) * len(clutter)
    max_total = n * (len(world) + stride)
    ctx.set_view(w, h, n, max_total)
    bufs = dict(views=views.nbytes, cq=cq.nbytes, lq=lq.nbytes, world=world.nbytes, cnt=4 * n, rec=32 * n * stride,
                cspr=64 * n * stride, spr=64 * max_total, off=4 * (n + 1), out=n * w * h * 4)
    d = {k: ctx.device_alloc(v) for k, v in bufs.items()}
    try:
        ctx.h2d(d["views"], views); ctx.h2d(d["cq"], cq); ctx.h2d(d["lq"], lq); ctx.h2d(d["world"], world)
        flags = L.PM3D_IN_DEVICE | L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC
        ctx._check(ctx.lib.pm3d_clutter_batch(ctx.h, d["cq"], n, cdist, stride, d["cnt"], d["rec"], d["cspr"], flags), "clutter")
        ctx._check(ctx.lib.pm3d_raycast_sprites_batch(ctx.h, d["world"], len(world), d["lq"], n, rd, d["cspr"], d["cnt"], stride,
                                                      None, None, max_total, d["spr"], d["off"], None, None, flags), "list")
        off = np.zeros(n + 1, dtype=np.int32)
        ctx.sync(); ctx.d2h(off, d["off"])               # the one host read: the total the render call is told
        ctx.render_resident(d["views"], n, d["spr"], d["off"], int(off[-1]), d["out"])
        frames = np.empty((n, h, w, 4), dtype=np.uint8)
        ctx.sync(); ctx.d2h(frames, d["out"])
    finally:
        for p in d.values():
            ctx.device_free(p)
    osc = O.Scene(scene)
    n_clutter_seen = 0
    for i, p in enumerate(poses):
        recs = O.clutter_update(scene.grid[0], cls, masks, defs.astype(O.CLUTTER_DEF_DTYPE), int(g["map_seed"]), cdist,
                                cq[i:i + 1].astype(O.CLUTTER_QUERY_DTYPE))
        csp = np.zeros(len(recs), dtype=O.SPRITE_DTYPE)
        csp["x"], csp["y"], csp["focusable"] = recs["x"], recs["y"], 1
        for k in ("scale", "sheet", "rx0", "ry0", "rx1", "ry1"):
            csp[k] = defs[k][recs["def"]]
        lst, _ = O.raycast_sprites(world.astype(O.SPRITE_DTYPE), p[0], p[1], rd, -1, csp)
        assert off[i + 1] - off[i] == len(lst)
        ref = O.render(osc, orc_camera(views[i], w, h), lst, threads=4)
        assert np.array_equal(frames[i], ref["rgba"]), f"view {i}"
        n_clutter_seen += len(recs)
    assert n_clutter_seen > 300


def _tick_setup(ctx, n, w, h, n_sets, rd=60.0, cdist=12.0):
    """Device-resident inputs of `n_sets` different pose sets for the chain clutter -> sprite list -> frames."""
    scene, g = W.proving_grounds_scene()
    scene.sheets = list(scene.sheets) + [W.sprite_sheet(18, 18, 41), W.sprite_sheet(24, 24, 42)]
    ctx.upload_scene(scene)
    clutter = [cl.MapClutter(str(im), float(f), float(hh), str(m) or None) for im, m, f, hh in
               zip(g["clutter_image"], g["clutter_match"], g["clutter_frequency"], g["clutter_height"])]
    sheets = {clutter[0].image: (len(scene.sheets) - 2, 18, 18), clutter[1].image: (len(scene.sheets) - 1, 24, 24)}
    defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(s) for s in g["floor_images"]], sheets)
    ctx.upload_clutter(defs, cls, masks, int(g["map_seed"]))
    world = W.golden_sprites(scene, g)
    stride = 21 * (int(cdist) + 1) * len(clutter)
    max_total = n * (len(world) + stride)
    ctx.set_view(w, h, n, max_total)
    poses = W.free_cell_poses(scene, n * n_sets, 2024)
    views = W.make_views(scene, w, h, poses, render_distance=rd)
    cq = cl.clutter_queries([(p[0], p[1], p[3]) for p in poses], 70.0)
    lq = sp.list_queries([(p[0], p[1]) for p in poses])
    bufs = dict(views=views.nbytes, cq=cq.nbytes, lq=lq.nbytes, world=world.nbytes, cnt=4 * n, rec=32 * n * stride,
                cspr=64 * n * stride, spr=64 * max_total, off=4 * (n + 1), out=n * w * h * 4)
    d = {k: ctx.device_alloc(v) for k, v in bufs.items()}
    ctx.h2d(d["views"], views); ctx.h2d(d["cq"], cq); ctx.h2d(d["lq"], lq); ctx.h2d(d["world"], world)
    return dict(d=d, n=n, stride=stride, max_total=max_total, n_world=len(world), rd=rd, cdist=cdist,
                sz=(views.itemsize, cq.itemsize, lq.itemsize))


def _run_ticks(ctx, s, ticks, n_sets, d_sums, synced):
    """`ticks` ticks; tick t uses pose set t % n_sets and leaves its per-view checksums at d_sums[t].  The render call
    is told the CAPACITY of the list (the host never reads the total of an unsynced tick)."""
    L = pm3d._lib
    d, n = s["d"], s["n"]
    flags = L.PM3D_IN_DEVICE | L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC
    vsz, csz, lsz = s["sz"]
    for t in range(ticks):
        k = (t % n_sets) * n
        ctx._check(ctx.lib.pm3d_clutter_batch(ctx.h, d["cq"] + k * csz, n, s["cdist"], s["stride"], d["cnt"], d["rec"], d["cspr"], flags), "clutter")
        if synced:
            ctx.sync()
        ctx._check(ctx.lib.pm3d_raycast_sprites_batch(ctx.h, d["world"], s["n_world"], d["lq"] + k * lsz, n, s["rd"], d["cspr"], d["cnt"],
                                                      s["stride"], None, None, s["max_total"], d["spr"], d["off"], None, None, flags), "list")
        if synced:
            ctx.sync()
        ctx.render_resident(d["views"] + k * vsz, n, d["spr"], d["off"], s["max_total"], d["out"])
        if synced:
            ctx.sync()
        ctx._check(ctx.lib.pm3d_frame_checksums(ctx.h, n, d_sums + t * n * 8, L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC), "checksums")
        if synced:
            ctx.sync()


@pytest.mark.parametrize("n,w,h", [(8, 320, 200), (64, 320, 200), (512, 160, 100)])
@pytest.mark.parametrize("lanes", [1, 3])
def test_unsynced_chain_equals_the_synced_chain(gpu_ctx, n, w, h, lanes):
    """200 ticks of poses -> clutter -> sprite list -> frames -> checksums with NO pm3d_sync between any two calls and
    poses that change every tick (consecutive ticks list different numbers of sprites, so a front kernel that ran
    beside the list kernels of its own tick would read counts instead of offsets).  Every tick's per-view checksums
    must equal those of the same chain run with a sync after every call.  Regression test of the round-1 N=8 crash:
    the frame's front stream did not wait for the ctx stream."""
    ctx = gpu_ctx
    ticks, n_sets = 200, 5
    ctx.set_option(pm3d._lib.OPT_LANES, lanes)
    s = _tick_setup(ctx, n, w, h, n_sets)
    d_a, d_b = ctx.device_alloc(ticks * n * 8), ctx.device_alloc(n_sets * n * 8)
    try:
        _run_ticks(ctx, s, ticks, n_sets, d_a, synced=False)
        ctx.sync()
        a = np.zeros((ticks, n), dtype=np.uint64)
        ctx.d2h(a, d_a)
        _run_ticks(ctx, s, n_sets, n_sets, d_b, synced=True)
        b = np.zeros((n_sets, n), dtype=np.uint64)
        ctx.d2h(b, d_b)
        off = np.zeros(n + 1, dtype=np.int32)
        ctx.d2h(off, s["d"]["off"])
    finally:
        for p in list(s["d"].values()) + [d_a, d_b]:
            ctx.device_free(p)
        ctx.set_option(pm3d._lib.OPT_LANES, 1)
        ctx.set_view(w, h, 1, 1)
    assert off[-1] > n * 50 and len(set(np.diff(off).tolist())) > 1   # the lists are not trivial
    assert len({tuple(r) for r in b.tolist()}) == n_sets               # ... and the pose sets give different frames
    for t in range(ticks):
        assert np.array_equal(a[t], b[t % n_sets]), f"tick {t}: {int((a[t] != b[t % n_sets]).sum())} of {n} views differ"


def test_resident_render_rejects_bad_offsets(gpu_ctx):
    """pm3d_render_resident cannot validate offsets that live in HBM: the view setup does, on the device — a view whose
    offsets are out of range renders without sprites and pm3d_sync reports PM3D_ERR_CAPACITY."""
    ctx = gpu_ctx
    cfg = W.config_c2(320, 200, 40, n_poses=1)
    ctx.upload_scene(cfg["scene"])
    ctx.set_view(320, 200, 1, 4096)
    views, sprites = cfg["views"], cfg["sprites"]
    big = np.resize(sprites, 4096)
    d_v, d_s, d_o, d_out = ctx.device_alloc(views.nbytes), ctx.device_alloc(big.nbytes), ctx.device_alloc(8), ctx.device_alloc(320 * 200 * 4)
    try:
        ctx.h2d(d_v, views); ctx.h2d(d_s, big)
        ctx.render(views, None)
        plain = ctx.frame_checksums(1)[0]
        for bad in ([0, 3000], [5, 2], [-4, 10], [4000, 4200]):   # > 2048 per view, not monotone, negative, past the total
            ctx.h2d(d_o, np.array(bad, dtype=np.int32))
            ctx.render_resident(d_v, 1, d_s, d_o, 4096, d_out)
            with pytest.raises(pm3d.Pm3dError, match="sprite_offset"):
                ctx.sync()
            assert ctx.frame_checksums(1)[0] == plain              # rendered, without sprites
            ctx.sync()                                              # the error was reported once
        ctx.h2d(d_o, np.array([0, 40], dtype=np.int32))
        ctx.render_resident(d_v, 1, d_s, d_o, 4096, d_out)
        ctx.sync()
        assert ctx.frame_checksums(1)[0] != plain
    finally:
        for p in (d_v, d_s, d_o, d_out):
            ctx.device_free(p)


def test_unsynced_list_overflow_is_reported(gpu_ctx):
    ctx = gpu_ctx
    L = pm3d._lib
    scene, g = W.proving_grounds_scene()
    world = W.golden_sprites(scene, g)
    q = sp.list_queries([(30.0, 30.0), (31.0, 29.0)])
    d_w, d_q = ctx.device_alloc(world.nbytes), ctx.device_alloc(q.nbytes)
    cap = len(world) + 3
    d_s, d_o = ctx.device_alloc(64 * cap), ctx.device_alloc(12)
    try:
        ctx.h2d(d_w, world); ctx.h2d(d_q, q)
        flags = L.PM3D_IN_DEVICE | L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC
        ctx._check(ctx.lib.pm3d_raycast_sprites_batch(ctx.h, d_w, len(world), d_q, 2, -1.0, None, None, 0, None, None, cap, d_s, d_o,
                                                      None, None, flags), "list")
        with pytest.raises(pm3d.Pm3dError, match="max_total"):
            ctx.sync()
        ctx.sync()
    finally:
        for p in (d_w, d_q, d_s, d_o):
            ctx.device_free(p)


# ---- recorded frame sequences (pm3d_frames_capture_begin / _end / pm3d_frames_replay)
@pytest.mark.parametrize("lanes,overlap", [(1, 1), (4, 1), (1, 0)])
def test_recorded_frame_sequence_replays_the_same_frames(gpu_ctx, lanes, overlap):
    """A ring of resident frames captured into one CUDA graph: every replay writes the frames the plain calls write
    (checksums per output slot), also after the poses were rewritten in place between two replays; other entry points
    are refused inside the capture; a reconfiguration retires the recording."""
    import ctypes as C

    cfg = W.config_c2(320, 180, 40, n_poses=6)
    w, h, ring = 320, 180, 6
    ctx = gpu_ctx
    ctx.upload_scene(cfg["scene"])
    ctx.set_option(pm3d._lib.OPT_OVERLAP, overlap)
    ctx.set_option(pm3d._lib.OPT_LANES, lanes)
    ctx.set_view(w, h, 1, 40)
    views, spr, off = cfg["views"], cfg["sprites"], cfg["offsets"]
    fb = w * h * 4
    dv, ds, do, dout = ctx.device_alloc(views.nbytes), ctx.device_alloc(spr.nbytes), ctx.device_alloc(8), ctx.device_alloc(ring * fb)
    ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, 40], np.int32))

    def frame(r):
        ctx.render_resident(dv + r * views.itemsize, 1, ds + int(off[r]) * spr.itemsize, do, 40, dout + r * fb)

    def ring_bytes():
        ctx.sync()
        out = np.zeros(ring * fb, dtype=np.uint8)
        ctx.d2h(out, dout)
        return out

    try:
        for r in range(ring):
            frame(r)
        plain = ring_bytes()
        ctx.h2d(dout, np.zeros(ring * fb, dtype=np.uint8))
        ctx.frames_capture_begin()
        for r in range(ring):
            frame(r)
        with pytest.raises(pm3d.Pm3dError):
            ctx.sync()                                   # refused inside a capture
        g = ctx.frames_capture_end()
        assert not ring_bytes().any()                    # capturing executed nothing
        for _ in range(3):
            ctx.frames_replay(g)
        assert np.array_equal(ring_bytes(), plain)
        # poses rewritten in place: the recording follows the data, not the values it saw
        ctx.h2d(dv, views[::-1].copy())
        ctx.frames_replay(g)
        swapped = ring_bytes().reshape(ring, fb)
        ctx.h2d(dv, views)
        assert not np.array_equal(swapped, plain.reshape(ring, fb))   # other cameras now
        ctx.frames_replay(g)
        frame(0)                                         # plain calls after a replay stay ordered behind it
        assert np.array_equal(ring_bytes(), plain)
        ctx.set_view(w, h, 2, 40)                        # reconfiguration retires the recording
        with pytest.raises(pm3d.Pm3dError):
            ctx.frames_replay(g)
    finally:
        ctx.set_option(pm3d._lib.OPT_LANES, 1)
        ctx.set_option(pm3d._lib.OPT_OVERLAP, 1)
        for p_ in (dv, ds, do, dout):
            ctx.device_free(p_)
