"""GPU parity of the full Game.getValidMove batch (pm3d_valid_move_batch, game/collision.go:69-355) against the
oracle's restatement: every field of every result bit for bit (same binary64 operation sequence, no contraction),
on bundled and synthetic maps, with the player block, sprite colliders, Z filters, long moves through many walls,
moves leaving the map and non-finite input.  The checkAlternate recursion of the host mirror (pm3d/collision.py)
is compared with the oracle's recursion as well."""
import math

import numpy as np
import pytest

import pm3d
from oracle import oracle as O
from pm3d import collision as col
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


def assert_results_equal(got, want):
    assert got.dtype == want.dtype and len(got) == len(want)
    for name in got.dtype.names:
        a, b = got[name], want[name]
        if a.dtype.kind == "f":
            a, b = a.view(np.uint64), b.view(np.uint64)
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, f"{name}: {len(bad)} of {len(got)} differ, first {bad[:5].tolist()}: gpu {got[bad[:3]]} oracle {want[bad[:3]]}"


def random_colliders(rng, level0, n, player=True):
    free = np.argwhere(level0 == 0)
    c = np.zeros(n, dtype=O.COLLIDER_DTYPE)
    cells = free[rng.integers(0, len(free), n)]
    c["x"] = cells[:, 0] + rng.random(n)
    c["y"] = cells[:, 1] + rng.random(n)
    c["z"] = np.where(rng.random(n) < 0.7, 0.0, rng.random(n) * 2.0)
    c["radius"] = np.where(rng.random(n) < 0.05, 0.0, 0.05 + rng.random(n) * 0.6)
    c["height"] = 0.1 + rng.random(n) * 1.5
    c["anchor"] = rng.integers(0, 3, n)
    c["destroyed"] = rng.random(n) < 0.05
    return c


def random_movers(rng, level0, colliders, n, max_len):
    free = np.argwhere(level0 == 0)
    m = np.zeros(n, dtype=O.MOVER_DTYPE)
    cells = free[rng.integers(0, len(free), n)]
    m["pos_x"] = cells[:, 0] + rng.random(n)
    m["pos_y"] = cells[:, 1] + rng.random(n)
    # a third of the movers start next to a collider so the circle stages are exercised
    near = rng.random(n) < 0.35
    if len(colliders):
        pick = rng.integers(0, len(colliders), n)
        m["pos_x"] = np.where(near, colliders["x"][pick] + rng.normal(0, 0.8, n), m["pos_x"])
        m["pos_y"] = np.where(near, colliders["y"][pick] + rng.normal(0, 0.8, n), m["pos_y"])
    m["pos_z"] = np.where(rng.random(n) < 0.5, 0.0, rng.random(n) * 1.5)
    ang = rng.random(n) * 2 * math.pi
    length = rng.random(n) ** 3 * max_len
    m["move_x"] = m["pos_x"] + np.cos(ang) * length
    m["move_y"] = m["pos_y"] + np.sin(ang) * length
    m["move_z"] = m["pos_z"] + np.where(rng.random(n) < 0.5, 0.0, rng.normal(0, 0.3, n))
    m["radius"] = np.where(rng.random(n) < 0.03, 0.0, 0.01 + rng.random(n) * 0.4)
    m["height"] = np.where(rng.random(n) < 0.03, 0.0, 0.01 + rng.random(n))
    nc = max(len(colliders), 1)
    m["self"] = np.where(rng.random(n) < 0.2, rng.integers(0, nc, n), -1)
    m["parent"] = np.where(rng.random(n) < 0.3, rng.integers(0, nc, n), -1)
    m["parent"] = np.where(rng.random(n) < 0.1, 0, m["parent"])   # the player's own projectiles
    m["anchor"] = rng.integers(0, 3, n)
    m["destroyed"] = rng.random(n) < 0.05
    noop = rng.random(n) < 0.02
    for a, b in (("move_x", "pos_x"), ("move_y", "pos_y"), ("move_z", "pos_z")):
        m[a] = np.where(noop, m[b], m[a])
    return m


@pytest.mark.parametrize("name,n_col,max_len", [("debug_land", 300, 3.0), ("proving_grounds", 408, 1.5), ("debug_land", 0, 40.0)])
def test_bundled_map_valid_move_parity(gpu_ctx, name, n_col, max_len):
    g = W.load_golden(name)
    rng = np.random.default_rng(7 + n_col)
    lvl = np.ascontiguousarray(g["levels"][:1])
    scene = W.scene_from_grid(lvl, 1)
    gpu_ctx.upload_scene(scene)
    gpu_ctx.upload_segments(g["segments"])
    c = random_colliders(rng, lvl[0], n_col)
    player = 0 if n_col else -1
    gpu_ctx.upload_colliders(c.astype(pm3d._lib.COLLIDER_DTYPE), player)
    m = random_movers(rng, lvl[0], c, 40000, max_len)
    got = gpu_ctx.valid_move_batch(m.astype(pm3d._lib.MOVER_DTYPE), W.CLIP_DISTANCE)
    want = O.valid_move_batch(O.Scene(scene), g["segments"], c, player, W.CLIP_DISTANCE, m, threads=8)
    assert_results_equal(got.view(O.MOVE_RESULT_DTYPE), want)
    assert 0 < int(want["collided"].sum()) < len(want)
    if n_col:
        assert int((want["n_collisions"] > 0).sum()) > 100       # the circle stages really ran
        assert int((want["first_entity"] == 0).sum()) > 0         # ... and so did the player block


def test_dense_map_long_moves_cross_many_walls(gpu_ctx):
    """256x256 at 10 % walls: moves of up to 120 cells cross dozens of wall outlines (more than the kernel's
    8-entry hit list: brute-force fallback) and leave the map."""
    grid = W.synthetic_levels(256, 256, 1, 0.10, 42)
    segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
    scene = W.scene_from_grid(grid, 1)
    rng = np.random.default_rng(99)
    gpu_ctx.upload_scene(scene)
    gpu_ctx.upload_segments(segs)
    c = random_colliders(rng, grid[0], 64)
    gpu_ctx.upload_colliders(c.astype(pm3d._lib.COLLIDER_DTYPE), 0)
    m = random_movers(rng, grid[0], c, 20000, 120.0)
    got = gpu_ctx.valid_move_batch(m.astype(pm3d._lib.MOVER_DTYPE), W.CLIP_DISTANCE)
    want = O.valid_move_batch(O.Scene(scene), segs, c, 0, W.CLIP_DISTANCE, m, threads=8)
    assert_results_equal(got.view(O.MOVE_RESULT_DTYPE), want)
    assert int((want["n_points"] > 8).sum()) > 50


def test_valid_move_edge_cases(gpu_ctx):
    lvl = np.zeros((1, 12, 10), dtype=np.uint16)
    lvl[0, 0, :] = lvl[0, -1, :] = 32767
    lvl[0, :, 0] = lvl[0, :, -1] = 32767
    lvl[0, 5, 3:7] = 1
    scene = W.scene_from_grid(lvl, 1)
    segs = pm3d.generate_wall_collision_lines(lvl[0], False, W.CLIP_DISTANCE)   # no inner border: moves can leave the map
    c = np.zeros(2, dtype=O.COLLIDER_DTYPE)
    c[0] = (3.5, 2.5, 0.0, 0.3, 1.0, 0, 0, [0] * 6)
    c[1] = (7.5, 6.5, 0.5, 0.4, 0.8, 1, 0, [0] * 6)
    rows = [
        ((2.5, 2.5, 0.0), (3.4, 2.5, 0.0)),         # circle hit
        ((2.5, 2.5, 0.0), (4.5, 2.5, 0.0)),         # through the player circle: repeated records
        ((3.3, 2.5, 0.0), (3.35, 2.5, 0.0)),        # inside the combined circle
        ((2.5, 4.5, 0.0), (8.5, 4.5, 0.0)),         # through the bar
        ((2.5, 2.5, 0.5), (2.75, 2.5, -0.25)),      # under the floor
        ((1.5, 1.5, 0.0), (-3.0, 1.5, 0.0)),        # leaves the map on -x
        ((10.5, 8.5, 0.0), (30.0, 30.0, 0.0)),      # leaves on +x, +y
        ((1.5, 1.5, 0.0), (1.5, 1.5, 0.0)),         # no-op
        ((2.5, 2.5, 0.0), (float("nan"), 2.5, 0.0)),
        ((2.5, 2.5, 0.0), (float("inf"), 2.5, 0.0)),
        ((2.5, 2.5, 0.0), (2.5, 2.5, 1.0)),         # Z only
        ((3.5, 2.5, 0.0), (3.6, 2.5, 0.0)),         # starts on the collider centre
        ((7.5, 6.0, 0.5), (7.5, 7.0, 0.5)),         # vertical move through collider 1 (B == 0 branch)
        ((7.0, 6.5, 0.5), (8.0, 6.5, 0.5)),         # tangent-ish / through its centre line
        ((7.1, 6.1, 0.5), (7.9, 6.9, 0.5)),
        ((5.0, 1.5, 0.0), (5.0, 8.5, 0.0)),         # along the bar's face line
        ((-50.0, -50.0, 0.0), (-49.0, -49.5, 0.0)),  # far outside the binned area
    ]
    m = np.zeros(len(rows) * 3, dtype=O.MOVER_DTYPE)
    for k in range(3):                              # every row with each anchor
        for i, (p, t) in enumerate(rows):
            m[k * len(rows) + i] = (*p, *t, 0.2, 0.5, -1, -1, k, 0, [0] * 6)
    gpu_ctx.upload_scene(scene)
    gpu_ctx.upload_segments(segs)
    for player in (-1, 0, 1):
        gpu_ctx.upload_colliders(c.astype(pm3d._lib.COLLIDER_DTYPE), player)
        got = gpu_ctx.valid_move_batch(m.astype(pm3d._lib.MOVER_DTYPE), W.CLIP_DISTANCE)
        want = O.valid_move_batch(O.Scene(scene), segs, c, player, W.CLIP_DISTANCE, m)
        assert_results_equal(got.view(O.MOVE_RESULT_DTYPE), want)
    with pytest.raises(pm3d.Pm3dError, match="player_index"):
        gpu_ctx.upload_colliders(c.astype(pm3d._lib.COLLIDER_DTYPE), 5)


def test_host_mirror_get_valid_move_with_check_alternate(gpu_ctx):
    """pm3d.collision.Game.getValidMove(entity, x, y, z, True): the recursion runs on the host, every geometric
    test on the device; the result equals the oracle's recursion (both take cos/sin/atan2 from the C library)."""
    g = W.load_golden("debug_land")
    lvl = np.ascontiguousarray(g["levels"][:1])
    scene = W.scene_from_grid(lvl, 1)
    gpu_ctx.upload_scene(scene)
    game = col.Game(gpu_ctx, g["segments"])
    rng = np.random.default_rng(5)
    c = random_colliders(rng, lvl[0], 120)
    ents = [col.Entity(float(r["x"]), float(r["y"]), float(r["z"]), float(r["radius"]), float(r["height"]), int(r["anchor"]),
                       bool(r["destroyed"])) for r in c]
    game.set_colliders(ents[0], ents[1:])
    osc = O.Scene(scene)
    m = random_movers(rng, lvl[0], c, 1500, 2.5)
    n_alt = 0
    for r in m:
        unit = col.Entity(float(r["pos_x"]), float(r["pos_y"]), float(r["pos_z"]), max(float(r["radius"]), 0.05),
                          max(float(r["height"]), 0.05), int(r["anchor"]), False,
                          heading=float(rng.random() * 2 * math.pi - math.pi), velocity=float(rng.random() * 0.2))
        pos, z, collided, hits = game.getValidMove(unit, float(r["move_x"]), float(r["move_y"]), float(r["move_z"]), True)
        rec = np.zeros(1, dtype=O.MOVER_DTYPE)
        rec[0] = (unit.x, unit.y, unit.z, r["move_x"], r["move_y"], r["move_z"], unit.collision_radius, unit.collision_height,
                  -1, -1, unit.anchor, 0, [0] * 6)
        want = O.valid_move_alternate(osc, g["segments"], c, 0, W.CLIP_DISTANCE, rec, unit.heading, unit.velocity, True)
        assert (pos[0], pos[1], z, int(collided)) == (want["x"], want["y"], want["z"], int(want["collided"])), (r, want)
        assert len(hits) == (1 if want["n_collisions"] > 0 else 0)
        if hits:
            assert hits[0].entity is game.colliders[int(want["first_entity"])]
            assert hits[0].collision == (want["first_x"], want["first_y"]) and hits[0].collisionZ == want["first_z"]
        n_alt += int(want["n_points"] > 0)
    assert n_alt > 30
    # lineOfSight through the same mirror
    assert game.lineOfSight(None, ents[0]) is False
    a, b = col.Entity(2.5, 2.5), col.Entity(3.5, 2.5)
    assert game.lineOfSight(a, b) == bool(O.los_batch(g["segments"], np.array([[2.5, 2.5, 3.5, 2.5]]))[0])


def test_binned_colliders_equal_the_full_loop_in_awkward_layouts(gpu_ctx):
    """The colliders are binned by cell when there are >= 32 of them: positions on cell borders, stacked in one cell
    (more candidates than the kernel's 16-entry list), outside the map, huge, infinite and NaN ("loose" list), movers
    landing exactly on borders and moves longer than the binned look-up handles (full-loop fallback)."""
    lvl = np.zeros((1, 24, 20), dtype=np.uint16)
    lvl[0, 0, :] = lvl[0, -1, :] = 32767
    lvl[0, :, 0] = lvl[0, :, -1] = 32767
    scene = W.scene_from_grid(lvl, 1)
    segs = pm3d.generate_wall_collision_lines(lvl[0], False, W.CLIP_DISTANCE)
    rng = np.random.default_rng(3)
    n = 96
    c = np.zeros(n, dtype=O.COLLIDER_DTYPE)
    c["x"] = rng.integers(1, 23, n) + rng.choice([0.0, 0.5, 0.999999999, 0.25], n)     # many exactly on borders
    c["y"] = rng.integers(1, 19, n) + rng.choice([0.0, 0.5, 0.75], n)
    c["x"][:24] = 10.0 + rng.random(24) * 0.9                                            # 24 stacked in one cell
    c["y"][:24] = 10.0 + rng.random(24) * 0.9
    c["x"][24:30] = [-3.5, 40.0, 1e7, float("inf"), float("nan"), -1e12]                 # loose
    c["y"][24:30] = [5.0, 5.0, 5.0, 5.0, 5.0, 5.0]
    c["radius"] = 0.05 + rng.random(n) * 0.5
    c["height"] = 0.2 + rng.random(n)
    c["anchor"] = rng.integers(0, 3, n)
    m = np.zeros(6000, dtype=O.MOVER_DTYPE)
    k = len(m)
    m["pos_x"] = rng.integers(1, 23, k) + rng.choice([0.0, 0.5, 0.3, 0.999], k)
    m["pos_y"] = rng.integers(1, 19, k) + rng.choice([0.0, 0.5, 0.7], k)
    m["pos_x"][:1500] = 9.0 + rng.random(1500) * 3.0                                     # around the stack
    m["pos_y"][:1500] = 9.0 + rng.random(1500) * 3.0
    ang = rng.random(k) * 2 * np.pi
    length = rng.choice([0.25, 0.5, 1.0, 2.5, 3.0, 7.0], k)                              # 7.0: wider than the binned look-up
    m["move_x"] = np.where(rng.random(k) < 0.3, np.round(m["pos_x"] + np.cos(ang) * length), m["pos_x"] + np.cos(ang) * length)
    m["move_y"] = np.where(rng.random(k) < 0.3, np.round(m["pos_y"] + np.sin(ang) * length), m["pos_y"] + np.sin(ang) * length)
    m["move_x"][-40:-20] = -2.5                                                          # next to the loose colliders
    m["move_y"][-40:-20] = 5.0 + rng.random(20) - 0.5
    m["move_x"][-20:] = 1e7 + rng.random(20)
    m["move_y"][-20:] = 5.0
    m["radius"], m["height"], m["self"], m["parent"], m["anchor"] = 0.2, 0.6, -1, -1, rng.integers(0, 3, k)
    gpu_ctx.upload_scene(scene)
    gpu_ctx.upload_segments(segs)
    for player in (-1, 3, 50):
        gpu_ctx.upload_colliders(c.astype(pm3d._lib.COLLIDER_DTYPE), player)
        got = gpu_ctx.valid_move_batch(m.astype(pm3d._lib.MOVER_DTYPE), W.CLIP_DISTANCE)
        want = O.valid_move_batch(O.Scene(scene), segs, c, player, W.CLIP_DISTANCE, m, threads=8)
        assert_results_equal(got.view(O.MOVE_RESULT_DTYPE), want)
    assert int((want["n_collisions"] > 16).sum()) > 10       # the stacked cell overflows the candidate list
