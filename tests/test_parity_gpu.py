"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the
same seeded inputs.

Bars (BASELINE.json north_star): DDA hit cells, wall sides and texture column indices
bit-exact; distances bit-exact (binary64, no FMA contraction on either side); final RGBA
pixels within <= 1 LSB on <= 0.1 % of pixels.  The integer blit model makes the pixels exact
as well, so these tests assert ZERO differing pixels and only fall back to the stated
tolerance in one place (test_frame_tolerance_statement) to document it.
"""
import math

import numpy as np
import pytest

import pm3d
from conftest import assert_columns_equal, orc_camera
from oracle import oracle as O
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


def render_both(ctx, scene, w, h, views, sprites, offsets=None, keep_columns=True):
    ctx.upload_scene(scene)
    ctx.set_option(pm3d._lib.OPT_KEEP_COLUMNS, 1 if keep_columns else 0)
    ctx.set_view(w, h, len(views), max(len(sprites), 1))
    if offsets is None:
        offsets = np.array([0, len(sprites)], dtype=np.int32)
    got = ctx.render(views, sprites, offsets)
    osc = O.Scene(scene)
    want = []
    for i, v in enumerate(views):
        want.append(O.render(osc, orc_camera(v, w, h), sprites[offsets[i]:offsets[i + 1]].astype(O.SPRITE_DTYPE), threads=4))
    return got, want


def assert_frame_equal(got_rgba, want_rgba, what=""):
    bad = (got_rgba != want_rgba).any(axis=-1)
    n = int(bad.sum())
    if n:
        ys, xs = np.nonzero(bad)
        first = [(int(x), int(y), got_rgba[y, x].tolist(), want_rgba[y, x].tolist()) for x, y in list(zip(xs, ys))[:5]]
        raise AssertionError(f"{what}: {n} of {bad.size} pixels differ, first (x, y, gpu, oracle): {first}")


def check_all(ctx, got, want, offsets, n_sprites):
    for i, ref in enumerate(want):
        cols, zbuf = ctx.download_columns(i)
        assert_columns_equal(cols, ref["columns"])
        assert np.array_equal(zbuf.view(np.uint64), ref["zbuf"].view(np.uint64)), "z-buffer bits"
        assert_frame_equal(got["rgba"][i], ref["rgba"], f"view {i}")
        lo, hi = int(offsets[i]), int(offsets[i + 1])
        assert np.array_equal(got["rects"][lo:hi], ref["rects"]), f"view {i}: sprite screen rects"
        c = got["conv"][i]
        assert int(c["sprite"]) == ref["convergence"]["sprite"]
        assert float(c["distance"]) == ref["convergence"]["distance"] or (math.isnan(float(c["distance"])) and math.isnan(ref["convergence"]["distance"]))
        assert (float(c["point_x"]), float(c["point_y"])) == ref["convergence"]["point"] or math.isnan(float(c["point_x"]))


# --------------------------------------------------------------------------- the five configs (reduced where the oracle would be slow)


@pytest.mark.parametrize("pixels", ["real", "synthetic"])
def test_c1_bundled_map_640x360(gpu_ctx, pixels):
    """pixels = "real": the reference's own wall / floor / sky PNGs and tree / mech sheets (tests/golden/
    proving_grounds_pixels.npz; the trees are mostly translucent texels, the boundary wall half holes);
    "synthetic": the procedural tiles (alpha-hole boundary tile, 1-px translucent rims)."""
    c = W.config_c1(pixels)
    off = np.array([0, len(c["sprites"])], dtype=np.int32)
    got, want = render_both(gpu_ctx, c["scene"], 640, 360, c["views"], c["sprites"], off)
    check_all(gpu_ctx, got, want, off, len(c["sprites"]))


@pytest.mark.parametrize("pixels", ["real", "synthetic"])
def test_c2_1080p_200_sprites(gpu_ctx, pixels):
    c = W.config_c2(n_poses=3, pixels=pixels)
    got, want = render_both(gpu_ctx, c["scene"], 1920, 1080, c["views"], c["sprites"], c["offsets"])
    check_all(gpu_ctx, got, want, c["offsets"], len(c["sprites"]))


@pytest.mark.parametrize("levels,pitch", [(3, 0.0), (8, math.pi / 4), (8, -math.pi / 8), (3, 0.2)])
def test_c3_multilevel_infinite_distance(gpu_ctx, levels, pitch):
    # the C3 scene at 960x540 (the 4K frame is covered by the checksum property test below)
    c = W.config_c3(levels=levels, w=960, h=540, pitch=pitch)
    off = np.array([0, 0], dtype=np.int32)
    got, want = render_both(gpu_ctx, c["scene"], 960, 540, c["views"], c["sprites"], off)
    check_all(gpu_ctx, got, want, off, 0)


@pytest.mark.parametrize("pixels", ["real", "synthetic"])
def test_c4_batched_views_320x200(gpu_ctx, pixels):
    c = W.config_c4(n_views=48, pixels=pixels)
    got, want = render_both(gpu_ctx, c["scene"], 320, 200, c["views"], c["sprites"], c["offsets"])
    check_all(gpu_ctx, got, want, c["offsets"], len(c["sprites"]))


# --------------------------------------------------------------------------- edge cases


def edge_scene():
    grid = W.synthetic_levels(24, 17, 3, 0.18, 77)
    return W.scene_from_grid(grid, 4)  # one raycast level more than stored: top level repeats


EDGE_POSES = [
    (5.5, 5.5, 0.5, 0.0, 0.0),                 # ray direction exactly +x: deltaDistY = +Inf
    (5.5, 5.5, 0.5, math.pi / 2, 0.0),         # exactly +y (cos = 6e-17)
    (6.0, 5.0, 0.5, 0.3, 0.0),                 # standing exactly on a cell corner
    (6.0, 5.5, 0.5, math.pi, 0.0),             # on a vertical cell border, looking along -x
    (5.5, 5.5, 0.05, 1.1, -math.pi / 8),       # low camera, looking down
    (5.5, 5.5, 2.9, 2.2, math.pi / 4),         # high camera (above level 2), max pitch up
    (0.5, 0.5, 0.5, 0.7, 0.0),                 # inside the boundary wall cell
    (-3.25, 8.5, 0.5, 0.0, 0.0),               # outside the map: first step leaves the grid (hit = 2)
    (30.0, 30.0, 0.5, 4.0, 0.0),               # far outside, looking away
    (12.5, 8.5, 0.5, 5.5, 0.1),
]


@pytest.mark.parametrize("rd", [-1.0, 6.0])
def test_edge_poses(gpu_ctx, rd):
    scene = edge_scene()
    w, h = 200, 120
    rng = W.SplitMix64(5)
    u = rng.uniform(60 * 4).reshape(60, 4)
    xyzs = np.stack([1 + 22 * u[:, 0], 1 + 15 * u[:, 1], 2.0 * u[:, 2] * (u[:, 3] > 0.5), 0.2 + u[:, 3]], axis=1)
    sheet = (u[:, 2] * 4).astype(np.int32)
    anchor = (u[:, 1] * 3).astype(np.uint8)
    spr = W.sprite_records(xyzs, sheet, anchor, scene.sheets, illumination=25.0)
    spr["focusable"][::3] = 0
    views = W.make_views(scene, w, h, EDGE_POSES, fov_deg=70.0, render_distance=rd)
    n = len(views)
    offsets = (np.arange(n + 1) * len(spr)).astype(np.int32)
    sprites = np.concatenate([spr] * n)
    got, want = render_both(gpu_ctx, scene, w, h, views, sprites, offsets)
    check_all(gpu_ctx, got, want, offsets, len(sprites))


def test_zoom_fov_depth_2(gpu_ctx):
    scene = edge_scene()
    views = W.make_views(scene, 256, 144, [(5.5, 5.5, 0.5, 0.4, 0.05)], fov_deg=35.0, fov_depth=2.0, render_distance=50.0)
    off = np.array([0, 0], dtype=np.int32)
    got, want = render_both(gpu_ctx, scene, 256, 144, views, np.zeros(0, pm3d._lib.SPRITE_DTYPE), off)
    check_all(gpu_ctx, got, want, off, 0)


@pytest.mark.parametrize("w,h", [(1, 1), (3, 2), (63, 31), (65, 33), (130, 9), (322, 5)])
def test_ragged_view_sizes(gpu_ctx, w, h):
    """Widths that are not multiples of 4 / 64 and tiny frames (scalar-store and edge-group paths)."""
    scene = edge_scene()
    views = W.make_views(scene, w, h, [(5.5, 5.5, 0.5, 0.9, 0.0)])
    xyzs = np.array([[8.5, 7.5, 0.0, 0.8], [7.0, 6.2, 0.3, 0.5]])
    spr = W.sprite_records(xyzs, np.array([0, 2], np.int32), np.array([0, 1], np.uint8), scene.sheets)
    off = np.array([0, 2], dtype=np.int32)
    got, want = render_both(gpu_ctx, scene, w, h, views, spr, off)
    check_all(gpu_ctx, got, want, off, 2)


def test_no_sprites_no_floor_no_boxes(gpu_ctx):
    """FloorTextureAt nil (floor texturing off, texture.go:120-122) and no sky / floor box."""
    scene = edge_scene()
    scene.floor_tiles = np.full_like(scene.floor_tiles, -1)
    scene.sky_tile = scene.floorbox_tile = -1
    views = W.make_views(scene, 160, 100, [(5.5, 5.5, 0.5, 0.9, 0.0)])
    off = np.array([0, 0], dtype=np.int32)
    got, want = render_both(gpu_ctx, scene, 160, 100, views, np.zeros(0, pm3d._lib.SPRITE_DTYPE), off)
    check_all(gpu_ctx, got, want, off, 0)
    assert (got["rgba"][0][..., 3] == 0).any()  # untouched pixels stay transparent


def test_many_overlapping_sprites_and_duplicates(gpu_ctx):
    """> 256 sprites in one band (band-list overflow path), exact duplicates (tie order), sprites behind
    the camera, zero scale, bad sheet ids and rects hanging off the sheet."""
    scene = edge_scene()
    w, h = 192, 108
    rng = W.SplitMix64(9)
    n = 400
    u = rng.uniform(n * 3).reshape(n, 3)
    xyzs = np.stack([7.0 + 3.0 * u[:, 0], 5.0 + 2.0 * u[:, 1], np.zeros(n), 0.3 + 0.5 * u[:, 2]], axis=1)
    xyzs[10:20] = xyzs[10]          # ten sprites at the same spot
    xyzs[20, :2] = (2.0, 5.5)       # behind the camera
    spr = W.sprite_records(xyzs, (u[:, 2] * 4).astype(np.int32), np.zeros(n, np.uint8), scene.sheets)
    spr["scale"][21] = 0.0
    spr["sheet"][22] = 99
    spr["sheet"][23] = -1
    spr["rx0"][24], spr["rx1"][24] = -20, 60   # rect partly outside the sheet
    spr["ry0"][25], spr["ry1"][25] = 150, 400
    spr["rx0"][26], spr["rx1"][26] = 50, 50    # empty rect
    views = W.make_views(scene, w, h, [(5.5, 5.5, 0.5, 0.0, 0.0)], render_distance=-1.0)
    off = np.array([0, n], dtype=np.int32)
    got, want = render_both(gpu_ctx, scene, w, h, views, spr, off)
    check_all(gpu_ctx, got, want, off, n)


@pytest.mark.parametrize("n", [257, 1024, 1500, 2048])
def test_view_setup_orderings(gpu_ctx, n):
    """The per-view sprite ordering has three implementations by list size — placement by counting (<= 256), rank sort
    (<= 1024) and the bitonic network (<= 2048, the per-view limit): sizes at and across the boundaries, with ties."""
    scene = edge_scene()
    w, h = 160, 90
    rng = W.SplitMix64(100 + n)
    u = rng.uniform(n * 3).reshape(n, 3)
    xyzs = np.stack([1.5 + 20.0 * u[:, 0], 1.5 + 13.0 * u[:, 1], np.zeros(n), 0.2 + 0.3 * u[:, 2]], axis=1)
    xyzs[5:40, :2] = np.round(xyzs[5:40, :2])   # many equal distances from the cell-centred camera
    spr = W.sprite_records(xyzs, (u[:, 2] * 4).astype(np.int32), np.zeros(n, np.uint8), scene.sheets)
    views = W.make_views(scene, w, h, [(12.5, 8.5, 0.5, 2.5, 0.0)], render_distance=-1.0)
    off = np.array([0, n], dtype=np.int32)
    got, want = render_both(gpu_ctx, scene, w, h, views, spr, off)
    check_all(gpu_ctx, got, want, off, n)


@pytest.mark.parametrize("rows", [2, 4, 8, 16])
def test_warp_list_overflow_in_both_sprite_modes(gpu_ctx, rows, monkeypatch):
    """700 billboards stacked on one patch overflow the warp-private sprite list in latency mode (64 entries, 2/4 rows
    per warp) and in throughput mode (256 entries, 8/16 rows per warp): the nearest entries are listed, the farther
    ones are reached through the group list.  Pixels, rectangles and convergence must equal the oracle's, and the
    literal painter loop must give the same bytes."""
    scene = edge_scene()
    w, h = 192, 108
    rng = W.SplitMix64(77)
    n = 700
    u = rng.uniform(n * 3).reshape(n, 3)
    xyzs = np.stack([7.0 + 3.0 * u[:, 0], 5.2 + 0.6 * u[:, 1], np.zeros(n), 0.3 + 0.5 * u[:, 2]], axis=1)
    spr = W.sprite_records(xyzs, (u[:, 2] * 4).astype(np.int32), np.zeros(n, np.uint8), scene.sheets)
    views = W.make_views(scene, w, h, [(5.5, 5.5, 0.5, 0.0, 0.0)], render_distance=-1.0)
    off = np.array([0, n], dtype=np.int32)
    monkeypatch.setenv("PM3D_COMPOSE_ROWS", str(rows))
    got, want = render_both(gpu_ctx, scene, w, h, views, spr, off)
    check_all(gpu_ctx, got, want, off, n)
    fast = got["rgba"].copy()
    gpu_ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 1)
    try:
        slow = gpu_ctx.render(views, spr, off)["rgba"].copy()
    finally:
        gpu_ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 0)
    assert np.array_equal(fast, slow)


@pytest.mark.parametrize("rows", [2, 4, 8, 16])
@pytest.mark.parametrize("case", ["plain", "floor_holes", "translucent_floor", "translucent_sky", "no_boxes", "odd_height", "ragged_width", "short_render_distance"])
def test_sky_and_floor_runs(gpu_ctx, rows, case, monkeypatch):
    """Every shape writes the sky-only / floor-only row pairs of a sprite-free warp in tight runs ahead of the general
    row loop (compose_runs: 2 columns x 2 rows per lane, row records by shuffle).  Whatever ends a run — nil floor tiles, a
    translucent floor or sky texel, rows beyond the render distance, a missing sky / floor box, a last row without a
    partner, a partial 64-column group — must hand the row pair to the general loop: frames equal the oracle's and
    the literal painter's, with sprites only on some of the views (a view's sprite rows take no runs)."""
    scene = edge_scene()
    w, h, rd = 256, 160, -1.0
    if case == "floor_holes":
        scene.floor_tiles = scene.floor_tiles.copy()
        scene.floor_tiles[::3, ::2] = -1
        scene.floor_tiles[7:11, :] = -1
    elif case == "translucent_floor":
        scene.atlas = scene.atlas.copy()
        ft = int(scene.floor_tiles[scene.floor_tiles >= 0][0])
        scene.atlas[ft, 10:40, 20:50, 3] = 128
        scene.atlas[ft, 100:110, :, 3] = 0
        scene.atlas[ft, 10:40, 20:50, :3] //= 2     # keep it premultiplied
        scene.atlas[ft, 100:110, :, :3] = 0
    elif case == "translucent_sky":
        scene.atlas = scene.atlas.copy()
        scene.atlas[scene.sky_tile, 5:60, 30:90, 3] = 200
        scene.atlas[scene.sky_tile, 5:60, 30:90, :3] = np.minimum(scene.atlas[scene.sky_tile, 5:60, 30:90, :3], 200)
    elif case == "no_boxes":
        scene.sky_tile = scene.floorbox_tile = -1
    elif case == "odd_height":
        h = 161
    elif case == "ragged_width":
        w = 300
    elif case == "short_render_distance":
        rd = 4.0
    poses = [(5.5, 5.5, 0.5, 0.9, 0.0), (12.5, 8.5, 0.5, 5.5, 0.1), (5.5, 5.5, 0.05, 1.1, -math.pi / 8),
             (5.5, 5.5, 2.9, 2.2, math.pi / 4), (20.5, 14.5, 0.5, 3.6, -0.2), (1.5, 1.5, 0.5, 0.8, 0.3)]
    views = W.make_views(scene, w, h, poses, render_distance=rd)
    xyzs = np.array([[8.5, 7.5, 0.0, 0.8], [7.0, 6.2, 0.3, 0.5], [14.0, 9.0, 0.0, 1.2]])
    spr = W.sprite_records(xyzs, np.array([0, 2, 1], np.int32), np.array([0, 1, 0], np.uint8), scene.sheets)
    counts = [3, 0, 3, 0, 0, 3]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    sprites = np.concatenate([spr if k else spr[:0] for k in counts])
    monkeypatch.setenv("PM3D_COMPOSE_ROWS", str(rows))
    got, want = render_both(gpu_ctx, scene, w, h, views, sprites, off, keep_columns=False)
    for i, ref in enumerate(want):
        assert_frame_equal(got["rgba"][i], ref["rgba"], f"{case} rows {rows} view {i}")
        assert int(got["conv"][i]["sprite"]) == ref["convergence"]["sprite"]
    fast = got["rgba"].copy()
    gpu_ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 1)
    try:
        slow = gpu_ctx.render(views, sprites, off)["rgba"].copy()
    finally:
        gpu_ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 0)
    assert np.array_equal(fast, slow)


@pytest.mark.parametrize("rows,minb", [(2, 6), (2, 8), (4, 6), (4, 8), (8, 6), (8, 8), (16, 8)])
def test_every_compose_instantiation_gives_the_same_frame(gpu_ctx, rows, minb, monkeypatch):
    """The launcher picks rows-per-warp, CTA shape and latency/throughput sprite stage by problem size; every
    instantiation must paint the oracle's frame (c2 scene at 640x360, 200 sprites, two poses)."""
    c = W.config_c2(w=640, h=360, n_poses=2)
    monkeypatch.setenv("PM3D_COMPOSE_ROWS", str(rows))
    monkeypatch.setenv("PM3D_COMPOSE_MINB", str(minb))
    got, want = render_both(gpu_ctx, c["scene"], 640, 360, c["views"], c["sprites"], c["offsets"], keep_columns=False)
    for i, ref in enumerate(want):
        assert_frame_equal(got["rgba"][i], ref["rgba"], f"rows {rows} minb {minb} view {i}")
        assert int(got["conv"][i]["sprite"]) == ref["convergence"]["sprite"]


def test_division_free_floor_weight_is_correctly_rounded(gpu_ctx):
    """The floor cast's weight = currentDist / distWall is evaluated as a Newton sequence on RN(1/distWall)
    (pm3d_internal.cuh div_by_rcp).  It must equal the IEEE quotient bit for bit: 2.1e9 operand pairs
    (random mantissas, all-ones / power-of-two / single-bit divisors, quotients beside rounding
    boundaries, the path's own value range, exponents to +-400)."""
    bad, n = gpu_ctx.selfcheck_division(20261019, 148 * 8, 7000)
    assert n > 2_000_000_000
    assert bad == 0, f"{bad} of {n} quotients differ from a / b"


def test_fast_floor_cast_index_equals_reference_sequence(gpu_ctx):
    """compose's fast floor cast (one FMA per coordinate + a rounding-boundary test, pm3d_internal.cuh) must give
    the reference's texel index whenever it accepts; 1.5e9 operand sets, half of them steered to within 2^14 ulps
    of an integer texel boundary."""
    bad, n, acc = gpu_ctx.selfcheck_floor(4242, 148 * 8, 5000)
    assert n > 1_000_000_000 and acc > n // 4
    assert bad == 0, f"{bad} of {acc} accepted fast floor casts differ from the reference sequence ({n} tested)"


def test_painter_and_front_to_back_paths_agree(gpu_ctx):
    """PM3D_OPT_FORCE_PAINTER runs the literal painter's loop on every pixel; the default
    front-to-back shortcut must give identical bytes (translucent boundary tile + sprite rims)."""
    c = W.config_c2(w=640, h=360, n_poses=2)
    gpu_ctx.upload_scene(c["scene"])
    gpu_ctx.set_view(640, 360, 2, len(c["sprites"]))
    fast = gpu_ctx.render(c["views"], c["sprites"], c["offsets"])["rgba"].copy()
    gpu_ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 1)
    try:
        slow = gpu_ctx.render(c["views"], c["sprites"], c["offsets"])["rgba"].copy()
    finally:
        gpu_ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 0)
    assert np.array_equal(fast, slow)
    assert ((fast[..., 3] > 0) & (fast[..., 3] < 255)).sum() == 0  # opaque sky/floor behind everything


def test_frame_tolerance_statement(gpu_ctx):
    """The stated tolerance: <= 1 LSB on <= 0.1 % of pixels.  (Measured: 0 pixels differ.)"""
    c = W.config_c2(w=960, h=540, n_poses=1)
    got, want = render_both(gpu_ctx, c["scene"], 960, 540, c["views"], c["sprites"], c["offsets"])
    diff = np.abs(got["rgba"][0].astype(np.int32) - want[0]["rgba"].astype(np.int32))
    assert diff.max() <= 1
    assert (diff.max(axis=-1) > 0).mean() <= 0.001


# --------------------------------------------------------------------------- API behaviour


@pytest.mark.parametrize("lanes", [2, 3, 4])
def test_frame_lanes_give_the_same_frames(gpu_ctx, lanes):
    """PM3D_OPT_LANES: independent resident frames dealt to several lanes (streams + intermediate buffers of their own)
    must be the frames a single lane renders; the readers of "the last render" follow the lane it ran on."""
    ctx = gpu_ctx
    c = W.config_c2(w=640, h=360, n_poses=7)
    views, sprites, off = c["views"], c["sprites"], c["offsets"]
    n, fb = len(views), 640 * 360 * 4
    ctx.upload_scene(c["scene"])
    ctx.set_view(640, 360, 1, 200)
    dv, ds, do = ctx.device_alloc(views.nbytes), ctx.device_alloc(sprites.nbytes), ctx.device_alloc(8)
    dout = ctx.device_alloc(3 * n * fb)
    ctx.h2d(dv, views); ctx.h2d(ds, sprites); ctx.h2d(do, np.array([0, 200], np.int32))

    def run(reps):
        outs = []
        for i in range(reps * n):
            r = i % n
            ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, 200, dout + i * fb)
        rects, conv = ctx.download_results(200, 1)          # of the LAST call, whichever lane it ran on
        sums = ctx.frame_checksums(1)
        ctx.sync()
        frames = np.empty((reps * n, 360, 640, 4), dtype=np.uint8)
        ctx.d2h(frames, dout)
        return frames, rects.copy(), conv.copy(), int(sums[0])

    try:
        want, wrects, wconv, wsum = run(1)
        ctx.set_option(pm3d._lib.OPT_LANES, lanes)
        ctx.timer_start(9)
        got, grects, gconv, gsum = run(3)
        ctx.timer_stop(9)
        assert ctx.timer_ms(9) > 0
        for i in range(3 * n):
            assert np.array_equal(got[i], want[i % n]), f"frame {i} (lane {i % lanes})"
        # 3 * n - 1 frames precede the last one: it is pose n - 1 again
        assert np.array_equal(grects, wrects) and gconv.tobytes() == wconv.tobytes() and gsum == wsum
        # the host-buffer path keeps working next to the lanes (lane 0)
        one = ctx.render(views[2:3], sprites[off[2]:off[3]])
        assert np.array_equal(one["rgba"][0], want[2])
    finally:
        ctx.set_option(pm3d._lib.OPT_LANES, 1)
        for p_ in (dv, ds, do, dout):
            ctx.device_free(p_)


def test_frame_lanes_stress_with_interleaved_readers(gpu_ctx):
    """400 resident frames on 4 lanes; every few frames a reader of "the last render" (checksum, rectangles) or a
    host-buffer render cuts in.  Every reading must be that of the pose just rendered."""
    ctx = gpu_ctx
    c = W.config_c2(w=320, h=180, n_sprites=40, n_poses=11)
    views, sprites, off = c["views"], c["sprites"], c["offsets"]
    n, fb = len(views), 320 * 180 * 4
    ctx.upload_scene(c["scene"])
    ctx.set_view(320, 180, 1, 40)
    want_sum, want_rects = [], []
    for i in range(n):
        r = ctx.render(views[i:i + 1], sprites[off[i]:off[i + 1]])
        want_sum.append(int(ctx.frame_checksums(1)[0]))
        want_rects.append(r["rects"].copy())
    dv, ds, do = ctx.device_alloc(views.nbytes), ctx.device_alloc(sprites.nbytes), ctx.device_alloc(8)
    dout = ctx.device_alloc(16 * fb)
    ctx.h2d(dv, views); ctx.h2d(ds, sprites); ctx.h2d(do, np.array([0, 40], np.int32))
    rng = W.SplitMix64(2024)
    pick = (rng.uniform(400) * n).astype(int)
    act = (rng.uniform(400) * 10).astype(int)
    ctx.set_option(pm3d._lib.OPT_LANES, 4)
    try:
        for i in range(400):
            r = int(pick[i])
            ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, 40, dout + (i % 16) * fb)
            if act[i] == 0:
                assert int(ctx.frame_checksums(1)[0]) == want_sum[r], f"frame {i}"
            elif act[i] == 1:
                rects, _ = ctx.download_results(40, 1)
                assert np.array_equal(rects, want_rects[r]), f"frame {i}"
            elif act[i] == 2:
                h = ctx.render(views[r:r + 1], sprites[off[r]:off[r + 1]])
                assert np.array_equal(h["rects"], want_rects[r]) and int(ctx.frame_checksums(1)[0]) == want_sum[r]
        ctx.sync()
    finally:
        ctx.set_option(pm3d._lib.OPT_LANES, 1)
        for p_ in (dv, ds, do, dout):
            ctx.device_free(p_)


def host_checksum(frame: np.ndarray) -> int:
    """The per-view checksum pm3d_frame_checksums defines: sum over words of FNV-1a style hash(word index, word)."""
    words = np.ascontiguousarray(frame).view(np.uint32).reshape(-1).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = (np.uint64(1469598103934665603) ^ np.arange(len(words), dtype=np.uint64)) * np.uint64(1099511628211)
        h = (h ^ words) * np.uint64(1099511628211)
        return int(h.sum(dtype=np.uint64))


@pytest.mark.parametrize("w,h,n", [(320, 200, 3), (64, 128, 2), (61, 7, 2), (1920, 1080, 1)])
def test_frame_checksum_is_the_documented_hash(gpu_ctx, w, h, n):
    """Whole chunks (vector loads), a ragged tail and frames smaller than one chunk."""
    scene = edge_scene()
    views = W.make_views(scene, w, h, [(5.5, 5.5, 0.5, 0.3 + 0.9 * i, 0.0) for i in range(n)])
    gpu_ctx.upload_scene(scene)
    gpu_ctx.set_view(w, h, n, 1)
    got = gpu_ctx.render(views, None, np.zeros(n + 1, dtype=np.int32))
    sums = gpu_ctx.frame_checksums(n)
    assert [int(v) for v in sums] == [host_checksum(got["rgba"][i]) for i in range(n)]


def test_batch_equals_single_and_resident_equals_host(gpu_ctx):
    c = W.config_c4(n_views=12)
    ctx = gpu_ctx
    ctx.upload_scene(c["scene"])
    ctx.set_view(320, 200, 12, len(c["sprites"]))
    batch = ctx.render(c["views"], c["sprites"], c["offsets"])
    sums = ctx.frame_checksums(12)
    off = c["offsets"]
    for i in (0, 5, 11):
        one = ctx.render(c["views"][i:i + 1], c["sprites"][off[i]:off[i + 1]])
        assert np.array_equal(one["rgba"][0], batch["rgba"][i])
        assert ctx.frame_checksums(1)[0] == sums[i]
    # device-resident entry point
    views, sprites = c["views"], c["sprites"]
    dv, ds, do = ctx.device_alloc(views.nbytes), ctx.device_alloc(sprites.nbytes), ctx.device_alloc(off.nbytes)
    dout = ctx.device_alloc(12 * 320 * 200 * 4)
    ctx.h2d(dv, views); ctx.h2d(ds, sprites); ctx.h2d(do, off)
    ctx.render_resident(dv, 12, ds, do, len(sprites), dout)
    back = np.empty((12, 200, 320, 4), dtype=np.uint8)
    ctx.d2h(back, dout)
    assert np.array_equal(back, batch["rgba"])
    rects, conv = ctx.download_results(len(sprites), 12)
    assert np.array_equal(rects, batch["rects"]) and np.array_equal(conv["sprite"], batch["conv"]["sprite"])
    assert np.array_equal(ctx.frame_checksums(12), sums)
    for p in (dv, ds, do, dout):
        ctx.device_free(p)


def test_full_size_4k_frame_properties(gpu_ctx):
    """C3 at its full 3840x2160: the oracle renders this frame too (a few seconds on 4 threads)."""
    c = W.config_c3(levels=8, pitch=0.15)
    off = np.array([0, 0], dtype=np.int32)
    got, want = render_both(gpu_ctx, c["scene"], 3840, 2160, c["views"], c["sprites"], off, keep_columns=True)
    check_all(gpu_ctx, got, want, off, 0)
    assert (got["rgba"][0][..., 3] == 255).all()


def test_full_size_c4_sweep_properties(gpu_ctx, monkeypatch):
    """C4 at its full size: 4096 views of 320x200 with 50 sprites each, frames kept in HBM.  Every 64th view's
    device checksum equals the checksum of the ORACLE's frame of that view; the literal painter loop and another
    compose instantiation give the same 4096 checksums."""
    c = W.config_c4()
    ctx = gpu_ctx
    n = len(c["views"])
    assert n == 4096
    ctx.upload_scene(c["scene"])
    ctx.set_view(320, 200, n, len(c["sprites"]))
    ctx.render(c["views"], c["sprites"], c["offsets"], want_frame=False)
    sums = ctx.frame_checksums(n).copy()
    osc, off = O.Scene(c["scene"]), c["offsets"]
    for i in range(0, n, 64):
        ref = O.render(osc, orc_camera(c["views"][i], 320, 200), c["sprites"][off[i]:off[i + 1]].astype(O.SPRITE_DTYPE), threads=4)
        assert int(sums[i]) == host_checksum(ref["rgba"]), f"view {i}"
    ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 1)
    try:
        ctx.render(c["views"], c["sprites"], c["offsets"], want_frame=False)
        assert np.array_equal(ctx.frame_checksums(n), sums)
    finally:
        ctx.set_option(pm3d._lib.OPT_FORCE_PAINTER, 0)
    monkeypatch.setenv("PM3D_COMPOSE_ROWS", "4")
    ctx.render(c["views"], c["sprites"], c["offsets"], want_frame=False)
    assert np.array_equal(ctx.frame_checksums(n), sums)
    ctx.set_view(320, 200, 1, 1)   # release the 4096-view buffers


def test_errors_are_reported_not_swallowed(gpu_ctx):
    ctx = pm3d.Context(0)
    v = np.zeros(1, dtype=pm3d._lib.VIEW_DTYPE)
    with pytest.raises(pm3d.Pm3dError, match="must be set first|must precede"):
        ctx.set_view(64, 64, 1, 0)
    scene = edge_scene()
    ctx.upload_scene(scene)
    ctx.set_view(64, 64, 2, 10)
    with pytest.raises(pm3d.Pm3dError, match="max_views"):
        ctx.render(np.zeros(3, dtype=pm3d._lib.VIEW_DTYPE), None, np.zeros(4, np.int32))
    with pytest.raises(pm3d.Pm3dError, match="max_sprites_total"):
        ctx.render(v, np.zeros(11, dtype=pm3d._lib.SPRITE_DTYPE))
    with pytest.raises(pm3d.Pm3dError, match="upload_segments"):
        ctx.los_batch(np.zeros((1, 4)))
    bad = edge_scene()
    bad.floor_tiles = bad.floor_tiles.copy()
    bad.floor_tiles[3, 3] = 99          # tile outside the atlas: refused, not read out of bounds
    ctx.upload_scene(bad)
    with pytest.raises(pm3d.Pm3dError, match="outside the"):
        ctx.render(v)
    bad = edge_scene()
    bad.grid = bad.grid.copy()
    bad.grid[0, 2, 2] = 40000           # texture index beyond math.MaxInt16
    with pytest.raises(pm3d.Pm3dError, match="32767"):
        ctx.upload_scene(bad)
    ctx.close()


def test_camera_mirror_drives_the_path(gpu_ctx):
    """The raycaster.Camera mirror: NewCamera + setters + Update/GetConvergenceSprite/Draw, with the
    map / texture handler / sprites supplied through the callback interfaces."""
    from pm3d import camera as cam

    g = W.load_golden("debug_land")
    atlas = W.standard_atlas()
    level_map = cam.LevelMap(g["levels"], int(g["num_raycast_levels"]))
    textures = {int(i): (atlas[int(i) % 4], None, None) for i in g["tex_index"] if int(i) > 0}
    tex = cam.TextureHandler(level_map, textures, atlas[7], g["floor_grid"], [atlas[7], atlas[8], atlas[8], atlas[8]])
    sheets = W.default_sheets()
    c = cam.Camera(400, 240, 256, level_map, tex, ctx=gpu_ctx, sheets=sheets)
    c.SetRenderDistance(100.0)
    c.SetAlwaysSetSpriteScreenRect(True)
    c.SetFloorTexture(atlas[10]); c.SetSkyTexture(atlas[9])
    c.SetLightFalloff(float(g["light_falloff"])); c.SetGlobalIllumination(float(g["light_illumination"]))
    c.SetLightRGB(tuple(g["min_light"]), tuple(g["max_light"]))
    c.SetPosition((50.5, 50.5)); c.SetPositionZ(0.5); c.SetHeadingAngle(math.pi / 4); c.SetPitchAngle(0.0)
    c.SetFovAngle(70.0, 1.0)
    sprites = [cam.Sprite(x, y, z, s, cam.ANCHOR_BOTTOM, sheets[0]) for x, y, z, s in g["sprite_xyzs"][:60]]
    c.Update(sprites)
    frame = c.Draw()
    assert frame.shape == (240, 400, 4)
    # same numbers through the oracle
    scene = cam.flatten_scene(level_map, tex, 256, atlas[9], atlas[10], sheets)
    snap = cam.snapshot_sprites(sprites, {id(s): i for i, s in enumerate(sheets)})
    ref = O.render(O.Scene(scene), orc_camera(c.state.view_record()[0], 400, 240), snap.astype(O.SPRITE_DTYPE))
    assert_frame_equal(frame, ref["rgba"], "camera mirror")
    conv = c.GetConvergenceSprite()
    assert (conv is None and ref["convergence"]["sprite"] < 0) or conv is sprites[ref["convergence"]["sprite"]]
    for s, rc in zip(sprites, ref["rects"]):
        assert s.screen_rect == (None if rc[0] > rc[2] else tuple(int(v) for v in rc))
    assert c.FovDepth() == 1.0 and c.FovRadians() == pytest.approx(math.radians(70))
    c.SetViewSize(200, 120)
    c.Update(sprites)
    assert c.Draw().shape == (120, 200, 4)


def test_pipelined_frames_equal_synchronous_frames(gpu_ctx):
    """PM3D_OPT_PIPELINE: NO_SYNC renders into host memory overlap the frame's D2H with the next
    frame's kernels; after pm3d_sync every frame, rect and convergence must equal the synchronous result."""
    c = W.config_c2(w=640, h=360, n_poses=7)
    ctx = gpu_ctx
    ctx.upload_scene(c["scene"])
    ctx.set_view(640, 360, 1, 200)
    off = c["offsets"]
    want = [ctx.render(c["views"][i:i + 1], c["sprites"][off[i]:off[i + 1]]) for i in range(7)]
    want = [{k: v.copy() for k, v in r.items()} for r in want]
    ctx.set_option(pm3d._lib.OPT_PIPELINE, 3)
    try:
        bufs = [ctx.host_alloc(640 * 360 * 4, np.uint8, (1, 360, 640, 4)) for _ in range(7)]
        got = [ctx.render(c["views"][i:i + 1], c["sprites"][off[i]:off[i + 1]], out=bufs[i], no_sync=True) for i in range(7)]
        ctx.sync()
        for i in range(7):
            assert np.array_equal(got[i]["rgba"], want[i]["rgba"]), f"frame {i}"
            assert np.array_equal(got[i]["rects"], want[i]["rects"])
            assert np.array_equal(got[i]["conv"]["sprite"], want[i]["conv"]["sprite"])
    finally:
        ctx.set_option(pm3d._lib.OPT_PIPELINE, 1)


def test_pipelined_small_results_survive_slot_reuse(gpu_ctx):
    """The rectangles and convergence of pipelined renders reach the host through a ring of 8 mapped slots that the
    library flushes into the caller's buffers at pm3d_sync or when a slot comes round again: 27 frames in flight
    between two syncs, every frame with result buffers of its own."""
    c = W.config_c2(w=320, h=180, n_poses=9)
    ctx = gpu_ctx
    ctx.upload_scene(c["scene"])
    ctx.set_view(320, 180, 1, 200)
    off = c["offsets"]
    want = [ctx.render(c["views"][i:i + 1], c["sprites"][off[i]:off[i + 1]]) for i in range(9)]
    want = [{k: v.copy() for k, v in r.items()} for r in want]
    ctx.set_option(pm3d._lib.OPT_PIPELINE, 3)
    ring, ctx._res_ring = ctx._res_ring, []          # every call gets fresh (pageable) rect / convergence arrays
    try:
        bufs = [ctx.host_alloc(320 * 180 * 4, np.uint8, (1, 180, 320, 4)) for _ in range(27)]
        got = [ctx.render(c["views"][i % 9:i % 9 + 1], c["sprites"][off[i % 9]:off[i % 9 + 1]], out=bufs[i], no_sync=True)
               for i in range(27)]
        ctx.sync()
        for i in range(27):
            w_ = want[i % 9]
            assert np.array_equal(got[i]["rgba"], w_["rgba"]), f"frame {i}"
            assert np.array_equal(got[i]["rects"], w_["rects"]), f"rects of frame {i}"
            assert got[i]["conv"].tobytes() == w_["conv"].tobytes(), f"convergence of frame {i}"
    finally:
        ctx._res_ring = ring
        ctx.set_option(pm3d._lib.OPT_PIPELINE, 1)


def test_overlapped_frames_equal_serial_frames(gpu_ctx):
    """PM3D_OPT_OVERLAP (default on): the cast/setup launch of frame i+1 runs beside the compose launch of frame i
    on a second stream with a second set of intermediate buffers.  Many frames enqueued back to back (resident
    inputs, and host inputs with PM3D_NO_SYNC) must give the bytes, rectangles and convergence of the serial path."""
    ctx = gpu_ctx
    n = 9
    c = W.config_c2(w=640, h=360, n_sprites=60, n_poses=n)
    off = c["offsets"]
    ctx.upload_scene(c["scene"])
    ctx.set_view(640, 360, 1, 60)
    fb = 640 * 360 * 4

    def run_resident():
        d_v, d_s, d_o, d_out = ctx.device_alloc(c["views"].nbytes), ctx.device_alloc(c["sprites"].nbytes), ctx.device_alloc(8), ctx.device_alloc(n * fb)
        ctx.h2d(d_v, c["views"]); ctx.h2d(d_s, c["sprites"]); ctx.h2d(d_o, np.array([0, 60], np.int32))
        for rep in range(3):                      # reuse each set several times
            for i in range(n):
                ctx.render_resident(d_v + i * 104, 1, d_s + int(off[i]) * 64, d_o, 60, d_out + i * fb)
        ctx.sync()
        frames = np.zeros((n, 360, 640, 4), np.uint8)
        ctx.d2h(frames, d_out)
        rects, conv = ctx.download_results(60, 1)
        for p in (d_v, d_s, d_o, d_out):
            ctx.device_free(p)
        return frames, rects.copy(), conv.copy()

    def run_host():
        bufs = [ctx.host_alloc(fb, np.uint8, (1, 360, 640, 4)) for _ in range(n)]
        got = [ctx.render(c["views"][i:i + 1], c["sprites"][off[i]:off[i + 1]], out=bufs[i], no_sync=True) for i in range(n)]
        ctx.sync()
        return got

    try:
        ctx.set_option(pm3d._lib.OPT_OVERLAP, 0)
        f0, r0, c0 = run_resident()
        h0 = [{k: np.copy(v) for k, v in g.items()} for g in run_host()]
        ctx.set_option(pm3d._lib.OPT_OVERLAP, 1)
        f1, r1, c1 = run_resident()
        h1 = run_host()
    finally:
        ctx.set_option(pm3d._lib.OPT_OVERLAP, 1)
    assert np.array_equal(f0, f1)
    assert np.array_equal(r0, r1) and c0.tobytes() == c1.tobytes()
    assert len({f0[i].tobytes() for i in range(n)}) == n        # the poses really differ
    for i in range(n):
        assert np.array_equal(h0[i]["rgba"], h1[i]["rgba"]) and np.array_equal(h0[i]["rgba"][0], f0[i]), f"frame {i}"
        assert np.array_equal(h0[i]["rects"], h1[i]["rects"]) and h0[i]["conv"].tobytes() == h1[i]["conv"].tobytes()


def test_sparse_upper_levels_skip_the_walk_without_changing_a_pixel(gpu_ctx):
    """proving_grounds has 12 and 3 wall cells on levels 1 and 2: rays of those levels that provably pass none of
    them skip the DDA walk (cast_kernel.cuh; only without PM3D_OPT_KEEP_COLUMNS).  Frames from poses next to,
    facing, grazing and facing away from those cells must equal the oracle's, which walks every ray."""
    scene, g = W.proving_grounds_scene()
    poses = [
        (14.5, 27.5, 0.5, 2.98, 0.0),     # looking at the level-1/2 blocks at x = 8..10
        (9.5, 27.5, 0.5, 1.5708, 0.05),   # between two of them, looking along the row
        (9.5, 31.5, 0.6, -1.5708, -0.1),  # right behind one
        (12.0, 30.5, 0.5, 3.14159, 0.2),  # facing one head on, close
        (12.0, 30.5, 0.5, 0.0, 0.0),      # same spot facing away
        (7.2, 24.2, 0.5, 0.7854, 0.0),    # grazing a block's corner
        (480.5, 488.5, 0.5, 0.15, 0.0),   # the far cluster at (490..492, 490)
        (250.5, 250.5, 0.5, 0.9, 0.0),    # mid map: every upper-level ray skips
        (31.5, 3.5, 0.5, 0.7854, 0.0),
    ]
    views = W.make_views(scene, 480, 270, poses)
    ctx = gpu_ctx
    ctx.upload_scene(scene)
    ctx.set_option(pm3d._lib.OPT_KEEP_COLUMNS, 0)
    ctx.set_view(480, 270, len(poses), 1)
    off = np.zeros(len(poses) + 1, dtype=np.int32)
    got = ctx.render(views, np.zeros(0, pm3d._lib.SPRITE_DTYPE), off)
    osc = O.Scene(scene)
    n_upper_walls = 0
    for i, v in enumerate(views):
        ref = O.render(osc, orc_camera(v, 480, 270), np.zeros(0, O.SPRITE_DTYPE), threads=4)
        assert_frame_equal(got["rgba"][i], ref["rgba"], f"pose {i}")
        n_upper_walls += int((ref["columns"]["tile"][1:] >= 0).sum())
    assert n_upper_walls > 200     # the upper-level walls are really on screen in some of these frames


# ---- closed-form DDA jumps (pm3d_internal.cuh dda_jump, cast_kernel.cuh jump mode)
def test_dda_jump_is_certified_exact(gpu_ctx):
    """The wall cast advances N steps of the reference's DDA in closed form whenever a certificate on the two
    rounded side-distance sequences holds.  37888 rays (random, exact diagonals, rational slopes, cell centres,
    lattice positions, (almost) axis-parallel incl. inf / NaN side distances) x ~100 random step counts up to
    2000 against the sequential loop: every certified state must be the loop's."""
    bad, n, cert = gpu_ctx.selfcheck_dda(20261020, 148, 2000)
    assert n > 3_000_000
    assert bad == 0, f"{bad} of {cert} certified jump states differ from the sequential DDA ({n} tested)"
    assert cert > 0.6 * n, f"only {cert} of {n} states certified: the jump would rarely be used"


def _columns_and_frames(ctx, c, offsets):
    got = ctx.render(c["views"], c["sprites"], offsets)
    cols = [ctx.download_columns(i) for i in range(len(c["views"]))]
    return got, cols


@pytest.mark.parametrize("which", ["c2", "c3", "edge", "graze"])
def test_jump_and_sequential_cast_agree(gpu_ctx, which):
    """PM3D_OPT_DDA_JUMP = 0 walks every ray with the sequential loop; the default jump mode must give the same
    column records (hit cell, side, distances bit for bit), z-buffer and pixels."""
    if which == "c2":
        c = W.config_c2(w=640, h=360, n_poses=6)
    elif which == "c3":
        c = W.config_c3(levels=8, w=480, h=270, pitch=0.2, map_size=128)
    elif which == "graze":
        # rays running along walls for hundreds of cells (the long jumps over wall-free bounding boxes): the boundary
        # ring, a long interior wall, a diagonal staircase, corridors one cell wide
        grid = np.zeros((2, 300, 260), dtype=np.uint16)
        grid[0, 0, :] = grid[0, -1, :] = W.BOUNDARY_TEX
        grid[0, :, 0] = grid[0, :, -1] = W.BOUNDARY_TEX
        grid[0, 40:260, 100] = 1
        grid[:, 20:280, 180] = 2
        grid[0, 20:280, 182] = 3
        for k in range(60):
            grid[0, 100 + k, 20 + k] = 4
        scene = W.scene_from_grid(grid, 3)
        poses = [(1.5, 1.5, 0.5, 0.001, 0.0), (1.2, 258.7, 0.5, -0.002, 0.0), (298.5, 1.5, 0.5, math.pi / 2 + 0.003, 0.0),
                 (30.5, 100.9, 0.5, 0.0005, 0.0), (41.0, 99.5, 0.5, 0.01, 0.0), (21.5, 181.5, 0.5, 0.0, 0.0),
                 (279.5, 181.5, 0.5, math.pi, 0.0), (98.5, 17.5, 0.5, math.pi / 4, 0.0), (99.5, 21.5, 0.5, math.pi / 4 + 0.004, 0.0),
                 (150.0, 130.0, 0.5, 2.0, 0.0)]
        c = dict(scene=scene, w=320, h=120, sprites=np.zeros(0, O.SPRITE_DTYPE),
                 views=W.make_views(scene, 320, 120, poses, fov_deg=70.0, render_distance=-1.0))
    else:
        scene = edge_scene()
        c = dict(scene=scene, w=200, h=120, sprites=np.zeros(0, O.SPRITE_DTYPE),
                 views=W.make_views(scene, 200, 120, EDGE_POSES + [(5.5, 5.5, 0.5, math.pi / 4, 0.0), (5.0, 5.0, 0.5, math.pi / 4, 0.0)],
                                    fov_deg=70.0, render_distance=-1.0))
    n = len(c["views"])
    offsets = c.get("offsets")
    if offsets is None:
        offsets = (np.arange(n + 1) * len(c["sprites"])).astype(np.int32)
        c["sprites"] = np.concatenate([c["sprites"]] * n) if len(c["sprites"]) else c["sprites"]
    gpu_ctx.upload_scene(c["scene"])
    gpu_ctx.set_option(pm3d._lib.OPT_KEEP_COLUMNS, 1)
    gpu_ctx.set_view(c["w"], c["h"], n, max(len(c["sprites"]), 1))
    try:
        gpu_ctx.set_option(pm3d._lib.OPT_DDA_JUMP, 0)
        seq, seq_cols = _columns_and_frames(gpu_ctx, c, offsets)
        gpu_ctx.set_option(pm3d._lib.OPT_DDA_JUMP, 1)
        jmp, jmp_cols = _columns_and_frames(gpu_ctx, c, offsets)
    finally:
        gpu_ctx.set_option(pm3d._lib.OPT_DDA_JUMP, 1)
    for (ca, za), (cb, zb) in zip(seq_cols, jmp_cols):
        for f in [f for f in ca.dtype.names if not f.startswith("_")]:  # (field by field: the records carry padding bytes)
            assert ca[f].tobytes() == cb[f].tobytes(), f
        assert za.tobytes() == zb.tobytes()
    assert np.array_equal(seq["rgba"], jmp["rgba"])
