"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), the host mirrors of the reference interfaces, and the
workload generators."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import pm3d
from oracle import mapgen
from pm3d import camera as cam
from pm3d import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(ensure_built):
    with open(os.path.join(ROOT, "include", "pm3d.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(pm3d_[a-z0-9_]+)\s*\(", header))
    declared -= {"pm3d_ctx"}
    assert len(declared) >= 30
    lib = ctypes.CDLL(pm3d.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"libpm3d.so lacks {missing}"
    # and the Python binding knows the same set
    assert declared == set(pm3d._lib.SYMBOLS), declared ^ set(pm3d._lib.SYMBOLS)
    assert pm3d.load().pm3d_version() == 100


def test_struct_sizes_match_the_header(ensure_built):
    assert pm3d._lib.VIEW_DTYPE.itemsize == 104
    assert pm3d._lib.SPRITE_DTYPE.itemsize == 64
    assert pm3d._lib.COLUMN_DTYPE.itemsize == 56
    assert pm3d._lib.CONV_DTYPE.itemsize == 32
    assert pm3d._lib.MOVER_DTYPE.itemsize == 80
    assert pm3d._lib.COLLIDER_DTYPE.itemsize == 48
    assert pm3d._lib.MOVE_RESULT_DTYPE.itemsize == 80
    assert pm3d._lib.CLUTTER_DEF_DTYPE.itemsize == 40
    assert pm3d._lib.CLUTTER_QUERY_DTYPE.itemsize == 360
    assert pm3d._lib.CLUTTER_SPRITE_DTYPE.itemsize == 32
    assert pm3d._lib.LIST_QUERY_DTYPE.itemsize == 24
    # ... and the header says the same (the sizes are written next to each struct) and C agrees
    with open(os.path.join(ROOT, "include", "pm3d.h")) as f:
        header = f.read()
    for name, size in (("pm3d_view", 104), ("pm3d_sprite", 64), ("pm3d_column", 56), ("pm3d_mover", 80), ("pm3d_collider", 48),
                       ("pm3d_move_result", 80), ("pm3d_clutter_def", 40), ("pm3d_clutter_query", 360), ("pm3d_clutter_sprite", 32),
                       ("pm3d_list_query", 24)):
        assert re.search(r"\}\s*%s;\s*/\*\s*%d bytes" % (name, size), header), name
    import subprocess, tempfile
    src = '#include "pm3d.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pm3d_view), sizeof(pm3d_sprite), sizeof(pm3d_column), sizeof(pm3d_convergence), sizeof(pm3d_mover), sizeof(pm3d_collider), sizeof(pm3d_move_result), sizeof(pm3d_rect), sizeof(pm3d_clutter_def), sizeof(pm3d_clutter_query), sizeof(pm3d_clutter_sprite), sizeof(pm3d_list_query));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "s.c"), "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        out = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    assert [int(x) for x in out] == [104, 64, 56, 32, 80, 48, 80, 16, 40, 360, 32, 24]


def test_no_cpu_fallback(ensure_built):
    """Without a CUDA device the product fails loudly instead of computing on the CPU."""
    if pm3d.load().pm3d_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pm3d.Pm3dError, match="no CPU fallback"):
        pm3d.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pixelmek-3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "pm3d_oracle" not in src, fn


# --------------------------------------------------------------------------- L0 host mirror


@pytest.mark.parametrize("name", ["proving_grounds", "debug_land"])
def test_wall_collision_lines_match_oracle_on_bundled_maps(ensure_built, name):
    g = W.load_golden(name)
    got = pm3d.generate_wall_collision_lines(g["levels"][0], bool(g["has_boundary_wall"]), W.CLIP_DISTANCE)
    assert np.array_equal(got, g["segments"])


@pytest.mark.parametrize("seed,density,boundary", [(1, 0.05, True), (2, 0.3, True), (3, 0.6, False), (4, 0.95, True), (5, 0.0, True)])
def test_wall_collision_lines_match_oracle_on_random_maps(ensure_built, seed, density, boundary):
    lvl = W.synthetic_levels(40, 33, 1, density, seed, boundary=boundary)[0]
    want = mapgen.generate_wall_collision_lines(lvl, boundary, 0.01)
    got = pm3d.generate_wall_collision_lines(lvl, boundary, 0.01)
    assert np.array_equal(got, want)


def test_wall_collision_lines_empty_map(ensure_built):
    assert len(pm3d.generate_wall_collision_lines(np.zeros((0, 0), np.uint16), False, 0.01)) == 0


# --------------------------------------------------------------------------- interface mirrors


def test_level_map_semantics():
    lv = np.zeros((2, 4, 5), dtype=np.uint16)
    lv[1, 2, 3] = 7
    m = cam.LevelMap(lv, 4)
    assert m.Size() == (4, 5) and m.NumLevels() == 4
    assert m.Level(3) is not None and m.Level(3)[2][3] == 7  # above the top: top level repeats (map.go:259-263)
    assert m.IsWallAt(2, 2, 3) and not m.IsWallAt(0, 2, 3)


def test_texture_handler_side_selection_and_floor():
    img, sx, sy, dflt, path = (np.zeros((4, 4, 4), np.uint8) + k for k in range(5))
    lv = np.zeros((1, 3, 3), dtype=np.uint16)
    lv[0, 1, 1] = 2
    lv[0, 2, 2] = 1
    t = cam.TextureHandler(cam.LevelMap(lv), {1: (img, None, None), 2: (img, sx, sy)}, dflt,
                           floor_grid=np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]]), floor_images=[dflt, path])
    assert t.TextureAt(1, 1, 0, 0) is sy       # side 0 -> SideY (map.go:108)
    assert t.TextureAt(1, 1, 0, 1) is sx       # side 1 -> SideX (map.go:110)
    assert t.TextureAt(2, 2, 0, 0) is img
    assert t.TextureAt(0, 0, 0, 0) is None and t.TextureAt(-1, 0, 0, 0) is None and t.TextureAt(3, 0, 0, 0) is None
    assert t.FloorTextureAt(1, 1) is path and t.FloorTextureAt(0, 0) is dflt
    assert t.FloorTextureAt(-1, 0) is None
    t.SetRenderFloorTex(False)
    assert t.FloorTextureAt(1, 1) is None
    with pytest.raises(IndexError):
        t.SetRenderFloorTex(True)
        t.FloorTextureAt(5, 0)  # the Go code panics past the edge (texture.go:125)


def test_flatten_scene_resolves_callbacks():
    tiles = [W.wall_tile(i) for i in range(3)]
    floor_a, floor_b, sky, box = W.floor_tile(1), W.floor_tile(2), W.gradient_tile((1, 2, 3), (4, 5, 6)), W.floor_tile(3)
    lv = np.zeros((2, 6, 6), dtype=np.uint16)
    lv[0, 1, 1] = 1
    lv[0, 2, 2] = 2
    lv[1, 2, 2] = 2
    fg = np.zeros((6, 6), dtype=np.int32)
    fg[3, :] = 1
    m = cam.LevelMap(lv, 3)
    t = cam.TextureHandler(m, {1: (tiles[0], None, None), 2: (tiles[0], tiles[1], tiles[2])}, floor_a, fg, [floor_a, floor_b])
    s = cam.flatten_scene(m, t, 256, sky, box)
    lut = {int(k): (int(a), int(b)) for k, a, b in zip(s.wall_tex_index, s.wall_tile_side0, s.wall_tile_side1)}
    assert np.array_equal(s.atlas[lut[1][0]], tiles[0]) and lut[1][0] == lut[1][1]
    assert np.array_equal(s.atlas[lut[2][0]], tiles[2]) and np.array_equal(s.atlas[lut[2][1]], tiles[1])
    assert np.array_equal(s.atlas[s.floor_tiles[3, 0]], floor_b) and np.array_equal(s.atlas[s.floor_tiles[0, 0]], floor_a)
    assert np.array_equal(s.atlas[s.sky_tile], sky) and np.array_equal(s.atlas[s.floorbox_tile], box)
    assert s.num_levels == 3 and s.grid.shape == (2, 6, 6)


def test_camera_state_lowering():
    st = cam.CameraState(640, 360)
    st.pos, st.pos_z, st.heading, st.fov_angle, st.fov_depth = (3.0, 4.0), 0.75, math.pi / 2, math.radians(70), 1.0
    p = st.lowered()
    assert abs(p["dir_x"]) < 1e-15 and p["dir_y"] == 1.0
    # looking north (+y), the right of the screen is east (+x)
    assert p["plane_x"] == pytest.approx(math.tan(math.radians(35))) and abs(p["plane_y"]) < 1e-15
    assert p["cam_z_px"] == 0.25 * 360
    st.pitch_angle = math.pi / 4
    assert st.lowered()["pitch_px"] in (359, 360)  # int(h*tan(pi/4)): tan(pi/4) is 0.9999999999999999 in libm
    st.pitch_angle = 1.3
    assert st.lowered()["pitch_px"] == 360        # clamp upper = int(h * fovDepth)
    st.pitch_angle = -math.pi / 8
    assert st.lowered()["pitch_px"] == int(360 * math.tan(-math.pi / 8))
    st.pitch_angle = -1.2
    assert st.lowered()["pitch_px"] == -180      # clamp lower = -h/2
    # zoom: SetFovAngle(fov/2, 2.0) doubles |dir| (game/input.go:767)
    st.fov_angle, st.fov_depth, st.pitch_angle = math.radians(35), 2.0, 0.0
    p = st.lowered()
    assert math.hypot(p["dir_x"], p["dir_y"]) == pytest.approx(2.0)


def test_sheet_slices_are_one_pixel_short_like_the_reference():
    rects = cam.sheet_slices(240, 42, 6, 1)
    assert rects[0] == (0, 0, 39, 41) and rects[5] == (200, 0, 239, 41)  # sprite.go:176


def test_premultiply_matches_go_color_model():
    px = np.array([[[200, 100, 50, 128], [255, 255, 255, 255], [10, 20, 30, 0], [255, 0, 7, 1]]], dtype=np.uint8)
    got = pm3d.premultiply(px)[0]
    def go(c, a):
        return ((c * 0x101) * (a * 0x101) // 0xFFFF) >> 8
    for i in range(4):
        a = int(px[0, i, 3])
        assert got[i].tolist() == [go(int(px[0, i, k]), a) for k in range(3)] + [a]


def test_cardinal_to_angle():
    assert W.cardinal_to_angle(0) == pytest.approx(math.pi / 2)    # north
    assert W.cardinal_to_angle(90) == pytest.approx(0.0)           # east
    assert W.cardinal_to_angle(45) == pytest.approx(math.pi / 4)   # trial_day drop zone heading


# --------------------------------------------------------------------------- workloads


def test_splitmix64_reference_values():
    r = W.SplitMix64(0)
    # first outputs of splitmix64 seeded with 0 (public reference implementation)
    assert [r.next_u64() for _ in range(3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    a = W.SplitMix64(99).uniform(5)
    b = W.SplitMix64(99)
    assert a.tolist() == [(b.next_u64() >> 11) / float(1 << 53) for _ in range(5)]


def test_configs_are_deterministic_and_well_formed():
    c1a, c1b = W.config_c1(), W.config_c1()
    assert np.array_equal(c1a["views"], c1b["views"]) and np.array_equal(c1a["sprites"], c1b["sprites"])
    assert len(c1a["sprites"]) == 404 + 4 and c1a["scene"].grid.shape == (3, 500, 500)
    c2 = W.config_c2(n_poses=3)
    assert len(c2["views"]) == 3 and c2["offsets"].tolist() == [0, 200, 400, 600]
    for p in c2["poses"]:
        assert c2["scene"].grid[0][int(p[0])][int(p[1])] == 0
    c4 = W.config_c4(n_views=16)
    assert len(c4["views"]) == 16 and np.all(np.diff(c4["offsets"]) == 50)
    c5 = W.config_c5(n_rays=1000)
    assert c5["rays"].shape == (1000, 4) and len(c5["segments"]) > 100
    g = c5["grid"][0]
    assert (g[c5["rays"][:, 0].astype(int), c5["rays"][:, 1].astype(int)] == 0).all()


def test_shard_helpers():
    from pm3d import shard

    assert shard.view_shard(10, 1, 4).tolist() == [1, 5, 9]
    seen = []
    for r in range(3):
        lo, hi = shard.ray_shard(11, r, 3)
        seen += list(range(lo, hi))
    assert seen == list(range(11))


def test_library_shard_arithmetic_equals_the_host_helpers(ensure_built):
    import ctypes as C

    from pm3d import shard
    lib = pm3d.load()
    for n, world in ((4096, 8), (37, 3), (5, 8), (0, 2), (16777216, 7)):
        for r in range(world):
            assert lib.pm3d_shard_view_count(n, r, world) == len(shard.view_shard(n, r, world))
            lo, hi = C.c_int64(), C.c_int64()
            assert lib.pm3d_shard_ray_range(n, r, world, C.byref(lo), C.byref(hi)) == 0
            assert (lo.value, hi.value) == shard.ray_shard(n, r, world)
    assert lib.pm3d_shard_view_count(10, 3, 3) == -1
