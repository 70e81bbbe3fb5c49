"""CPU tests of the ORACLE itself: known answers, the committed fixtures and internal
consistency.  The reference ships no tests or golden vectors, so these
known answers are derived by hand from the algorithm and from facts SURVEY.md verified
independently (wall-cell counts computed from the YAML: 2095 / 12 / 3)."""
import math

import numpy as np
import pytest

from conftest import orc_camera
from oracle import mapgen
from oracle import oracle as O
from pm3d import workloads as W


# --------------------------------------------------------------------------- integer model


def test_go_int_matches_amd64_cvttsd2sq():
    assert O.go_int(3.99) == 3
    assert O.go_int(-3.99) == -3          # truncation toward zero
    assert O.go_int(float("nan")) == -(2 ** 63)
    assert O.go_int(float("inf")) == -(2 ** 63)
    assert O.go_int(-float("inf")) == -(2 ** 63)
    assert O.go_int(9.3e18) == -(2 ** 63)   # out of range -> "integer indefinite"
    assert O.go_int(-9.223372036854775808e18) == -(2 ** 63)
    assert O.go_int(9.2e18) == 9200000000000000000


# --------------------------------------------------------------------------- geom.LineIntersection


@pytest.mark.parametrize(
    "l1,l2,want",
    [
        ((0, 0, 2, 2), (0, 2, 2, 0), (1.0, 1.0, True)),      # plain crossing
        ((0, 0, 1, 0), (1, 0, 1, 1), (1.0, 0.0, True)),      # touching at an endpoint: inclusive
        ((0, 0, 1, 0), (0, 1, 1, 1), (0.0, 0.0, False)),     # parallel: denom == 0
        ((0, 0, 1, 0), (2, 0, 3, 0), (0.0, 0.0, False)),     # collinear: denom == 0 -> false
        ((0, 0, 1, 1), (2, 0, 2, 5), (0.0, 0.0, False)),     # t > 1
        ((0, 0, 4, 0), (2, -1, 2, -0.5), (0.0, 0.0, False)),  # u > 1 (segment stops short)
        ((0, 0, 4, 0), (2, -1, 2, 0), (2.0, 0.0, True)),     # u == 1 exactly
    ],
)
def test_line_intersection_known_answers(l1, l2, want):
    assert O.line_intersection(l1, l2) == want


# --------------------------------------------------------------------------- fixtures


def test_proving_grounds_fixture_matches_independent_counts():
    g = W.load_golden("proving_grounds")
    lv = g["levels"]
    assert lv.shape == (3, 500, 500)
    # counts computed from proving_grounds.yaml:31-66 independently
    assert [int((lv[i] > 0).sum()) for i in range(3)] == [2095, 12, 3]
    # boundary ring on level 0 only, value math.MaxInt16 (map.go:452-468)
    # (two generated wall polylines start ON the ring and overwrite cells (0,60) and (30,0) with texture 4)
    assert (lv[0, 0, :] > 0).all() and (lv[0, :, 499] == 32767).all() and (lv[0, 499, :] == 32767).all()
    assert int((lv[0] == 32767).sum()) == 4 * 499 - 2 and lv[0, 0, 60] == 4 and lv[0, 30, 0] == 4
    assert not (lv[1] == 32767).any()
    # 2 explicit + 300 filled + 98 stamped static sprites (proving_grounds.yaml:86-160)
    assert len(g["sprite_xyzs"]) == 400
    assert int(g["num_raycast_levels"]) == 3


def test_prefab_orientation():
    """PlusTower at [8,24] (proving_grounds.yaml): layers reversed in level AND y (map.go:491-503)."""
    lv = W.load_golden("proving_grounds")["levels"]
    # yaml layer 0 (top) = rows [0,0,0],[4,4,4],[0,0,0] -> level 1; layer 1 = [0,2,0],[1,1,1],[0,2,0] -> level 0
    assert lv[0, 8:11, 24:27].T.tolist() == [[0, 2, 0], [1, 1, 1], [0, 2, 0]][::-1]
    assert lv[1, 8:11, 24:27].T.tolist() == [[0, 0, 0], [4, 4, 4], [0, 0, 0]][::-1]
    assert (lv[2, 8:11, 24:27] == 0).all()


def test_wall_line_walk():
    """Wall polyline [[30,0],[30,50],[25,55]] with texture 4 on level 0 (map.go:527-549)."""
    lv = W.load_golden("proving_grounds")["levels"]
    assert (lv[0, 30, 1:51] == 4).all()
    assert lv[0, 27, 53] == 4 or lv[0, 27, 52] == 4  # the diagonal passes through here
    assert lv[1, 30, 10] == 0                      # height 1 only


def test_collision_lines_small_known_answer():
    """One isolated wall cell -> 4 border rect lines + 4 cell edges, clockwise from its west edge."""
    lvl = np.zeros((5, 5), dtype=np.uint16)
    lvl[2, 2] = 1
    lines = mapgen.generate_wall_collision_lines(lvl, False, 0.01)
    assert len(lines) == 8
    assert lines[:4].tolist() == [[0.01, 0.01, 0.01, 4.99], [0.01, 4.99, 4.99, 4.99], [4.99, 4.99, 4.99, 0.01], [4.99, 0.01, 0.01, 0.01]]
    assert lines[4:].tolist() == [[2, 2, 2, 3], [2, 3, 3, 3], [3, 3, 3, 2], [3, 2, 2, 2]]


def test_collision_lines_merge_runs():
    """A 3x1 bar of wall cells: interior edges cancel, collinear runs merge (map.go:783-809)."""
    lvl = np.zeros((6, 6), dtype=np.uint16)
    lvl[1:4, 2] = 1
    lines = mapgen.generate_wall_collision_lines(lvl, False, 0.0)[4:]
    assert lines.tolist() == [[1, 2, 1, 3], [1, 3, 4, 3], [4, 3, 4, 2], [4, 2, 1, 2]]


def test_golden_segments_reproduce():
    for name in ("proving_grounds", "debug_land"):
        g = W.load_golden(name)
        seg = mapgen.generate_wall_collision_lines(g["levels"][0], bool(g["has_boundary_wall"]), 0.01)
        assert np.array_equal(seg, g["segments"])


# --------------------------------------------------------------------------- wall cast known answers


def tiny_scene(levels=1):
    grid = np.zeros((levels, 8, 8), dtype=np.uint16)
    grid[0, 0, :] = grid[0, 7, :] = grid[0, :, 0] = grid[0, :, 7] = 1
    grid[0, 5, 3] = 2
    return W.scene_from_grid(grid, levels, has_boundary=False)


def test_dda_axis_aligned_known_answer():
    """Camera at (2.5, 3.5) looking along +x: the centre column hits cell (5,3) on its x side at
    perpendicular distance 2.5, line height h/2.5, texture column from wallX = 0.5."""
    scene = tiny_scene()
    w, h = 64, 40
    views = W.make_views(scene, w, h, [(2.5, 3.5, 0.5, 0.0, 0.0)], fov_deg=60.0, render_distance=-1.0)
    cols, zbuf = O.cast_walls(O.Scene(scene), orc_camera(views[0], w, h))
    c = cols[0][w // 2]
    assert (c["map_x"], c["map_y"], c["side"], c["hit"]) == (5, 3, 0, 1)
    assert c["perp_dist"] == 2.5
    assert c["line_height"] == int(h / 2.5)
    # texX = int(0.5*256) = 128, mirrored because side == 0 and rayDirX > 0: 256-128-1
    assert c["tex_x"] == 127
    assert c["draw_start"] == -(c["line_height"]) // 2 + h // 2 or c["draw_start"] == -(c["line_height"] // 2) + h // 2
    assert c["draw_end"] - c["draw_start"] == c["line_height"]
    assert zbuf[w // 2] == 2.5
    # texture index 2 has side variants: side 0 -> SideY image (tile 5) per map.go:108
    assert c["tile"] == 5
    # side-0 tint: 255-12 + sqrt(2.5)*(-100) + 300 clamped to max 255
    assert c["tint"].tolist()[:3] == [255, 255, 255]


def test_level_above_stored_repeats_top_level():
    scene = tiny_scene(levels=1)
    scene.num_levels = 3  # Map.Level(l >= len) returns the top level (map.go:259-263)
    w, h = 32, 20
    views = W.make_views(scene, w, h, [(2.5, 3.5, 0.5, 0.0, 0.0)], render_distance=-1.0)
    cols, _ = O.cast_walls(O.Scene(scene), orc_camera(views[0], w, h))
    for k in ("map_x", "map_y", "side", "tex_x"):
        assert np.array_equal(cols[0][k], cols[1][k]) and np.array_equal(cols[0][k], cols[2][k])
    lh = cols[0]["line_height"]
    assert np.array_equal(cols[2]["draw_start"], cols[0]["draw_start"] - 2 * lh)


def test_render_distance_culls_slice_but_keeps_zbuffer():
    scene = tiny_scene()
    w, h = 32, 20
    v_far = W.make_views(scene, w, h, [(2.5, 3.5, 0.5, 0.0, 0.0)], render_distance=1.0)
    cols, zbuf = O.cast_walls(O.Scene(scene), orc_camera(v_far[0], w, h))
    assert cols[0][w // 2]["tile"] == -1 and zbuf[w // 2] == 2.5


def test_floor_cast_known_answer():
    """Pixel (cx, y) of the floor: currentDist = h / (2y - h) for camZ=0, pitch=0; the floor point is
    pos + currentDist * rayDir, so looking along +x from (2.5,3.5) the centre column samples
    x = 2.5 + currentDist."""
    scene = tiny_scene()
    w, h = 64, 40
    views = W.make_views(scene, w, h, [(2.5, 3.5, 0.5, 0.0, 0.0)], fov_deg=60.0, render_distance=-1.0)
    res = O.render(O.Scene(scene), orc_camera(views[0], w, h), np.zeros(0, O.SPRITE_DTYPE))
    y = 35
    cd = h / (2.0 * y - h)
    fx, fy = 2.5 + cd * 1.0, 3.5
    tx, ty = int(fx * 256) % 256, int(fy * 256) % 256
    texel = scene.atlas[7][ty][tx]  # floor tile 7 everywhere
    tint = min(255, max(0, int(255 + math.sqrt(cd) * scene.light_falloff + scene.global_illumination)))
    want = [(int(texel[c]) * tint) >> 8 for c in range(3)] + [255]
    assert res["rgba"][y, w // 2].tolist() == want


def test_threads_do_not_change_the_result():
    c1 = W.config_c1()
    sc = O.Scene(c1["scene"])
    cam = orc_camera(c1["views"][0], 320, 180)
    spr = c1["sprites"].astype(O.SPRITE_DTYPE)
    a = O.render(sc, cam, spr, threads=1)
    b = O.render(sc, cam, spr, threads=4)
    assert np.array_equal(a["rgba"], b["rgba"]) and np.array_equal(a["rects"], b["rects"])


def test_sprite_order_far_to_near_stable():
    scene = tiny_scene()
    views = W.make_views(scene, 32, 20, [(2.5, 3.5, 0.5, 0.0, 0.0)])
    spr = np.zeros(4, dtype=O.SPRITE_DTYPE)
    spr["x"] = [3.5, 5.5, 3.5, 4.5]
    spr["y"] = 3.5
    order = O.sprite_order(orc_camera(views[0], 32, 20), spr)
    assert order.tolist() == [1, 3, 0, 2]  # farthest first; the tie (0, 2) keeps list order


def test_sprite_screen_rect_and_convergence():
    """A bottom-anchored unit sprite 2 cells straight ahead: height h/2, centred, feet on the
    floor line of a wall at the same distance."""
    scene = tiny_scene()
    w, h = 64, 40
    views = W.make_views(scene, w, h, [(2.5, 3.5, 0.5, 0.0, 0.0)], fov_deg=60.0, render_distance=-1.0)
    spr = np.zeros(1, dtype=O.SPRITE_DTYPE)
    spr["x"], spr["y"], spr["scale"], spr["sheet"] = 4.5, 3.5, 1.0, 3  # 18x18 sheet
    spr["rx1"], spr["ry1"], spr["focusable"] = 18, 18, 1
    res = O.render(O.Scene(scene), orc_camera(views[0], w, h), spr)
    x0, y0, x1, y1 = res["rects"][0].tolist()
    assert (y1 - y0) == h // 2 and (x1 - x0) == h // 2   # square texture, scale 1, depth 2
    assert (x0 + x1) // 2 == w // 2
    assert y1 == h // 2 + (h // 2) // 2                   # bottom = horizon + lineHeight(2)/2
    assert res["convergence"]["sprite"] == 0 and res["convergence"]["distance"] == 2.0


# --------------------------------------------------------------------------- LOS / move


def test_los_known_answers():
    lvl = np.zeros((8, 8), dtype=np.uint16)
    lvl[4, 2:6] = 1
    segs = mapgen.generate_wall_collision_lines(lvl, False, 0.01)
    rays = np.array([
        [1.5, 3.5, 6.5, 3.5],   # straight through the wall
        [1.5, 1.5, 6.5, 1.5],   # passes below it
        [1.5, 3.5, 3.5, 3.5],   # stops before it
        [1.5, 3.5, 4.0, 3.5],   # ends exactly on the wall's west face: endpoints inclusive -> blocked
        [2.5, 2.5, 2.5, 2.5],   # zero-length ray: denom == 0 everywhere -> visible
    ])
    assert O.los_batch(segs, rays).tolist() == [0, 1, 1, 0, 1]


def test_move_walls_known_answers():
    scene = tiny_scene()
    segs = mapgen.generate_wall_collision_lines(scene.grid[0], False, 0.01)
    moves = np.array([
        [2.5, 3.5, 3.5, 3.5],   # free move
        [2.5, 3.5, 5.5, 3.5],   # into the wall cell (5,3): crosses its outline -> stays
        [2.5, 3.5, 2.5, 3.5],   # no-op
    ])
    pos, col, hits = O.move_walls_batch(O.Scene(scene), segs, moves, 0.01)
    assert pos.tolist() == [[3.5, 3.5], [2.5, 3.5], [2.5, 3.5]]
    assert col.tolist() == [0, 1, 0]
    assert hits[1] >= 1


def test_go_golden_c1():
    """The oracle against the REAL Go raycaster at C1 — hit cells and wall sides of every column and level, sprite
    rectangles and the convergence record as dumped by baseline/go_harness (TextureAt / SetScreenRect recorders around
    raycaster-go's Camera.Update).  Skipped until tests/golden/go_c1.npz exists: no Go toolchain in this image."""
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "go_c1.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/go_c1.npz absent (baseline/go_harness needs a Go toolchain): parity stays unpinned")
    z = np.load(path)
    c1 = W.config_c1()
    sc = O.Scene(c1["scene"])
    cam = orc_camera(c1["views"][0], int(z["w"]), int(z["h"]))
    n_static = len(z["sprite_rects"])
    res = O.render(sc, cam, c1["sprites"][:n_static].astype(O.SPRITE_DTYPE), threads=4)
    cols = res["columns"]
    tex = z["texture_at"]
    for lv in range(int(z["levels"])):
        calls = tex[tex[:, 2] == lv]
        hit = cols["tile"][lv] >= 0   # the library asks for a texture where a wall was hit inside the map
        assert len(calls) >= int(hit.sum())
        got = np.stack([cols["map_x"][lv][hit], cols["map_y"][lv][hit], cols["side"][lv][hit]], axis=1)
        assert np.array_equal(calls[:len(got), [0, 1, 3]], got), f"level {lv}: hit cells / sides differ from the Go run"
    assert np.array_equal(res["rects"], z["sprite_rects"])
    assert res["convergence"]["distance"] == float(z["convergence"][0])
    assert res["convergence"]["sprite"] == int(z["convergence_sprite"])
