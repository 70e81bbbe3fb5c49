"""GPU parity tests of the clutter batch (pm3d_clutter_batch, csrc/clutter_kernel.cu) against the oracle's literal
restatement of ClutterHandler.Update + Go's math/rand.  Everything is compared bit for bit: counts, creation order,
cells, clutter indices and the binary64 sprite positions."""
import math

import numpy as np
import pytest

import pm3d
from oracle import oracle as O
from pm3d import clutter as cl
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


def oracle_batch(level0, cls, masks, defs, seed, dist, queries):
    return [O.clutter_update(level0, cls, masks, defs.astype(O.CLUTTER_DEF_DTYPE), seed, dist, queries[i:i + 1].astype(O.CLUTTER_QUERY_DTYPE))
            for i in range(len(queries))]


def assert_same(counts, recs, want, sprites=None, defs=None):
    for i, w in enumerate(want):
        assert int(counts[i]) == len(w), f"camera {i}: {int(counts[i])} sprites, oracle {len(w)}"
        g = recs[i, :len(w)]
        for k in ("cell_x", "cell_y", "def"):
            assert np.array_equal(g[k], w[k]), f"camera {i}: {k}"
        for k in ("x", "y"):
            assert np.array_equal(g[k].view(np.uint64), w[k].view(np.uint64)), f"camera {i}: {k} bits"
        if sprites is not None:
            s = sprites[i, :len(w)]
            assert np.array_equal(s["x"], w["x"]) and np.array_equal(s["y"], w["y"])
            assert not s["z"].any() and not s["illumination"].any() and (s["focusable"] == 1).all() and (s["anchor"] == 0).all()
            for k in ("scale", "sheet", "rx0", "ry0", "rx1", "ry1"):
                assert np.array_equal(s[k], defs[k][w["def"]])


def fixture_clutter(g):
    """Map.Clutter of the bundled map, frozen in tests/golden by make_fixtures.py"""
    return [cl.MapClutter(str(im), float(f), float(h), str(m) or None)
            for im, m, f, h in zip(g["clutter_image"], g["clutter_match"], g["clutter_frequency"], g["clutter_height"])]


def proving_grounds_clutter(ctx):
    scene, g = W.proving_grounds_scene()
    ctx.upload_scene(scene)
    clutter = fixture_clutter(g)
    assert [(c.image, c.frequency, c.height, c.floorPathMatch) for c in clutter] == [
        ("rocks/rock_0.png", 0.3, 1.0, "grass"), ("shrubbery/bush_0.png", 0.5, 1.0, "grass")]   # proving_grounds.yaml:77-85
    sheets = {c.image: (3, int(w), int(h)) for c, (w, h) in zip(clutter, g["clutter_image_size"])}
    defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(s) for s in g["floor_images"]], sheets)
    assert int(g["map_seed"]) == 12345                        # proving_grounds.yaml:5
    ctx.upload_clutter(defs, cls, masks, int(g["map_seed"]))
    return scene, g, defs, cls, masks


def test_proving_grounds_default_settings(gpu_ctx):
    """The bundled map with its own clutter list and seed, fov 70, clutterDistance 500 m = 25 cells (config.go:87,145)."""
    scene, g, defs, cls, masks = proving_grounds_clutter(gpu_ctx)
    poses = W.free_cell_poses(scene, 96, 5)
    cams = [(p[0], p[1], p[3]) for p in poses] + [(31.5, 3.5, W.cardinal_to_angle(45.0)), (0.5, 0.5, 0.0), (499.9, 250.0, 0.0)]
    q = cl.clutter_queries(cams, 70.0)
    counts, recs, sprites = gpu_ctx.clutter_batch(q, 25.0, 21 * 26 * 2, want_sprites=True)
    want = oracle_batch(scene.grid[0], cls, masks, defs, 12345, 25.0, q)
    assert_same(counts, recs, want, sprites, defs)
    assert counts.sum() > 10000 and counts.max() <= 21 * 26 * 2
    on_path = [r for w in want for r in w if g["floor_grid"][r["cell_x"], r["cell_y"]] != 0]
    assert not on_path                                        # floorPathMatch "grass": nothing on the tech floor


@pytest.mark.parametrize("mw,mh,dist,fov,n_defs", [(40, 40, 25.0, 70.0, 64), (23, 61, 12.5, 90.0, 5), (61, 23, 7.0, 45.0, 17),
                                                    (300, 200, 190.0, 110.0, 3), (16, 16, 0.5, 179.0, 1), (9, 120, 60.0, 20.0, 33)])
def test_random_maps_many_definitions(gpu_ctx, mw, mh, dist, fov, n_defs):
    """Non-square maps (posId = (x-1)*mapWidth + y collides on maps higher than wide: the first visit wins), up to 64
    clutter types with random floor masks and frequencies, cameras on the border and outside the map."""
    rng = W.SplitMix64(mw * 1000 + mh)
    grid = W.synthetic_levels(mw, mh, 1, 0.12, mw + mh)
    scene = W.scene_from_grid(grid, 1)
    gpu_ctx.upload_scene(scene)
    defs = np.zeros(n_defs, dtype=pm3d._lib.CLUTTER_DEF_DTYPE)
    defs["frequency"] = rng.uniform(n_defs) * 1.2 - 0.1
    defs["scale"] = 0.02 + rng.uniform(n_defs)
    defs["sheet"] = np.arange(n_defs) % 4
    defs["rx1"], defs["ry1"] = 17, 23
    n_cls = 7
    cls = (rng.uniform(mw * mh) * n_cls).astype(np.uint8).reshape(mw, mh)
    masks = np.array([rng.next_u64() for _ in range(n_cls)], dtype=np.uint64)
    masks[0] = 0                                              # a floor nothing grows on
    masks[1] = (1 << 64) - 1
    seed = [12345, -987654321987, 0, (1 << 63) - 5, -(1 << 63), 2147483647][(mw * 7 + mh) % 6]
    gpu_ctx.upload_clutter(defs, cls, masks, seed)
    u = rng.uniform(3 * 40).reshape(40, 3)
    cams = [(-3 + (mw + 6) * a, -3 + (mh + 6) * b, 2 * math.pi * c) for a, b, c in u]
    cams += [(0.0, 0.0, 0.0), (mw - 1e-9, mh / 2, math.pi), (mw / 2, -0.999, math.pi / 2), (1.5, 1.5, 5.0)]
    q = cl.clutter_queries(cams, fov)
    cap = 21 * (int(dist) + 1) * n_defs
    counts, recs = gpu_ctx.clutter_batch(q, dist, cap)
    want = oracle_batch(grid[0], cls, masks, defs, seed, dist, q)
    assert_same(counts, recs, want)
    assert counts.sum() > 0


def test_seed_that_normalises_to_zero(gpu_ctx):
    """Map.Seed + posId == 0 (mod 2^31-1) seeds with 89482311."""
    grid = np.zeros((1, 12, 12), dtype=np.uint16)
    gpu_ctx.upload_scene(W.scene_from_grid(grid, 1))
    defs = np.zeros(3, dtype=pm3d._lib.CLUTTER_DEF_DTYPE)
    defs["frequency"] = 1.0
    cls = np.zeros((12, 12), dtype=np.uint8)
    masks = np.array([0b111], dtype=np.uint64)
    pos_id = (5 - 1) * 12 + 6
    for seed in (-pos_id, 2147483647 - pos_id, -2147483647 * 3 - pos_id):
        gpu_ctx.upload_clutter(defs, cls, masks, seed)
        q = cl.clutter_queries([(5.5, 6.5, 0.0)], 70.0)
        counts, recs = gpu_ctx.clutter_batch(q, 3.0, 400)
        want = oracle_batch(grid[0], cls, masks, defs, seed, 3.0, q)
        assert_same(counts, recs, want)
        r = recs[0, 0]
        assert (r["cell_x"], r["cell_y"]) == (5, 6)
        g = O.GoRand(89482311)
        g.float64()
        assert r["x"] == 5 + g.float64()


def test_capacity_and_early_returns(gpu_ctx):
    scene, g, defs, cls, masks = proving_grounds_clutter(gpu_ctx)
    q = cl.clutter_queries([(100.5, 100.5, 1.0), (200.5, 300.5, 4.0)], 70.0)
    full_counts, full = gpu_ctx.clutter_batch(q, 25.0, 21 * 26 * 2)
    counts, recs = gpu_ctx.clutter_batch(q, 25.0, 50)          # fewer slots than sprites: full count, first 50 kept
    assert np.array_equal(counts, full_counts) and counts.min() > 50
    assert np.array_equal(recs, full[:, :50])
    for dist in (0.0, -3.0):                                   # clutter.go:32-34
        counts, _ = gpu_ctx.clutter_batch(q, dist, 10)
        assert not counts.any()
    gpu_ctx.upload_clutter(defs[:0], cls, masks, 12345)        # no clutter on the map
    counts, _ = gpu_ctx.clutter_batch(q, 25.0, 10)
    assert not counts.any()
    gpu_ctx.upload_clutter(defs, cls, masks, 12345)
    with pytest.raises(pm3d.Pm3dError):
        gpu_ctx.clutter_batch(q, 191.0, 10)
    with pytest.raises(pm3d.Pm3dError):
        gpu_ctx.clutter_batch(q, float("nan"), 10)


def test_handler_feeds_the_camera(gpu_ctx):
    """ClutterHandler.Update -> getRaycastSprites -> Camera.Update/Draw: the frame with the handler's sprites appended
    equals the oracle's frame with the oracle's clutter."""
    scene, g = W.proving_grounds_scene()
    rock = W.sprite_sheet(18, 18, 41)
    bush = W.sprite_sheet(24, 24, 42)
    scene.sheets = list(scene.sheets) + [rock, bush]
    i_rock, i_bush = len(scene.sheets) - 2, len(scene.sheets) - 1
    gpu_ctx.upload_scene(scene)
    clutter = fixture_clutter(g)
    handler = cl.ClutterHandler(gpu_ctx, clutter, g["floor_grid"], [str(s) for s in g["floor_images"]],
                                {"rocks/rock_0.png": (i_rock, 18, 18), "shrubbery/bush_0.png": (i_bush, 24, 24)}, 12345, 500)
    pose = (31.5, 3.5, 0.5, W.cardinal_to_angle(45.0), 0.0)
    handler.Update(pose[0], pose[1], pose[3], 70.0, 25.0, forceUpdate=True)
    assert len(handler.sprites) > 100 and len(handler.spritesByPosition) > 100
    kept = handler.sprites.copy()
    handler.Update(0.0, 0.0, 0.0, 70.0, 25.0, moved=False)     # camera did not move: nothing changes (clutter.go:26-29)
    assert np.array_equal(kept, handler.sprites)
    entity = W.golden_sprites(scene, g)
    sprites = np.concatenate([entity, handler.sprites])
    w, h = 480, 270
    views = W.make_views(scene, w, h, [pose])
    off = np.array([0, len(sprites)], dtype=np.int32)
    gpu_ctx.set_view(w, h, 1, len(sprites))
    got = gpu_ctx.render(views, sprites, off)
    # the oracle side, from the oracle's own clutter
    defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(s) for s in g["floor_images"]],
                                          {"rocks/rock_0.png": (i_rock, 18, 18), "shrubbery/bush_0.png": (i_bush, 24, 24)})
    q = cl.clutter_queries([(pose[0], pose[1], pose[3])], 70.0)
    recs = O.clutter_update(scene.grid[0], cls, masks, defs.astype(O.CLUTTER_DEF_DTYPE), 12345, 25.0, q.astype(O.CLUTTER_QUERY_DTYPE))
    osp = np.zeros(len(recs), dtype=O.SPRITE_DTYPE)
    osp["x"], osp["y"], osp["focusable"] = recs["x"], recs["y"], 1
    for k in ("scale", "sheet", "rx0", "ry0", "rx1", "ry1"):
        osp[k] = defs[k][recs["def"]]
    from conftest import orc_camera
    ref = O.render(O.Scene(scene), orc_camera(views[0], w, h), np.concatenate([entity.astype(O.SPRITE_DTYPE), osp]), threads=4)
    assert np.array_equal(got["rgba"][0], ref["rgba"])
    assert np.array_equal(got["rects"], ref["rects"])
    assert (got["rects"][len(entity):, 0] <= got["rects"][len(entity):, 2]).any()   # some clutter is on screen


def test_device_resident_batch(gpu_ctx):
    """PM3D_IN_DEVICE | PM3D_OUT_DEVICE | PM3D_NO_SYNC: queries and results stay in HBM."""
    scene, g, defs, cls, masks = proving_grounds_clutter(gpu_ctx)
    poses = W.free_cell_poses(scene, 256, 77)
    q = cl.clutter_queries([(p[0], p[1], p[3]) for p in poses], 70.0)
    n, cap = len(q), 21 * 26 * 2
    L = pm3d._lib
    d_q = gpu_ctx.device_alloc(q.nbytes)
    d_cnt = gpu_ctx.device_alloc(4 * n)
    d_rec = gpu_ctx.device_alloc(32 * n * cap)
    d_spr = gpu_ctx.device_alloc(64 * n * cap)
    try:
        gpu_ctx.h2d(d_q, q)
        rc = gpu_ctx.lib.pm3d_clutter_batch(gpu_ctx.h, d_q, n, 25.0, cap, d_cnt, d_rec, d_spr,
                                            L.PM3D_IN_DEVICE | L.PM3D_OUT_DEVICE | L.PM3D_NO_SYNC)
        assert rc == 0
        gpu_ctx.sync()
        counts = np.zeros(n, dtype=np.int32)
        recs = np.zeros((n, cap), dtype=L.CLUTTER_SPRITE_DTYPE)
        sprites = np.zeros((n, cap), dtype=L.SPRITE_DTYPE)
        gpu_ctx.d2h(counts, d_cnt); gpu_ctx.d2h(recs, d_rec); gpu_ctx.d2h(sprites, d_spr)
    finally:
        for p in (d_q, d_cnt, d_rec, d_spr):
            gpu_ctx.device_free(p)
    c2, r2, s2 = gpu_ctx.clutter_batch(q, 25.0, cap, want_sprites=True)
    assert np.array_equal(counts, c2)
    for i in range(n):
        assert np.array_equal(recs[i, :counts[i]], r2[i, :counts[i]]) and np.array_equal(sprites[i, :counts[i]], s2[i, :counts[i]])
