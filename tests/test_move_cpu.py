"""CPU tests of the full Game.getValidMove restatement (oracle/pm3d_oracle_move.c) against hand-derived
answers, and of the host-side pieces of pm3d/collision.py that need no GPU."""
import math

import numpy as np

from oracle import oracle as O
from pm3d import collision as col
from pm3d import workloads as W
import pm3d


def small_scene():
    lvl = np.zeros((1, 12, 10), dtype=np.uint16)
    lvl[0, 0, :] = lvl[0, -1, :] = 32767
    lvl[0, :, 0] = lvl[0, :, -1] = 32767
    lvl[0, 5, 3:7] = 1
    scene = W.scene_from_grid(lvl, 1)
    segs = pm3d.generate_wall_collision_lines(lvl[0], True, W.CLIP_DISTANCE)
    return scene, segs


def mover(pos, move, radius=0.2, height=0.5, self_i=-1, parent=-1, anchor=0, destroyed=0):
    m = np.zeros(1, dtype=O.MOVER_DTYPE)
    m[0] = (*pos, *move, radius, height, self_i, parent, anchor, destroyed, [0] * 6)
    return m


def colliders(*rows):
    c = np.zeros(len(rows), dtype=O.COLLIDER_DTYPE)
    for i, (x, y, z, r, h, anchor, destroyed) in enumerate(rows):
        c[i] = (x, y, z, r, h, anchor, destroyed, [0] * 6)
    return c


def test_line_circle_intersection_known_answers():
    # horizontal chord through the centre: larger x first ((-b + d) / 2a), both ends inside the segment
    assert O.line_circle_intersection([-2, 0, 2, 0], 0, 0, 1) == [(1.0, 0.0), (-1.0, 0.0)]
    # vertical line: B == 0 -> solved for y
    assert O.line_circle_intersection([0, -2, 0, 2], 0, 0, 1) == [(0.0, 1.0), (0.0, -1.0)]
    # the segment stops short of the circle: the infinite line intersects, the segment does not
    assert O.line_circle_intersection([2, 0, 3, 0], 0, 0, 1) == []
    assert len(O.line_circle_intersection([2, 0, 3, 0], 0, 0, 1, is_segment=False)) == 2
    # segment that ends inside the circle: one point
    assert O.line_circle_intersection([-2, 0, 0, 0], 0, 0, 1) == [(-1.0, 0.0)]
    # tangent: discriminant exactly 0 -> one point
    pts = O.line_circle_intersection([-2, 1, 2, 1], 0, 0, 1)
    assert len(pts) == 1 and pts[0][0] == 0.0 and pts[0][1] == 1.0
    # miss
    assert O.line_circle_intersection([-2, 1.5, 2, 1.5], 0, 0, 1) == []


def test_valid_move_free_and_noop_and_degenerate_entity():
    scene, segs = small_scene()
    s = O.Scene(scene)
    none = colliders()
    r = O.valid_move_batch(s, segs, none, -1, 0.01, mover((2.5, 2.5, 0.0), (3.25, 2.75, 0.0)))[0]
    assert (r["x"], r["y"], r["z"], r["collided"], r["n_collisions"], r["n_points"]) == (3.25, 2.75, 0.0, 0, 0, 0)
    assert r["first_entity"] == -1 and math.isnan(r["near_x"])
    # no-op move (collision.go:79-81) and an entity without a collision volume (:76-78) return the position untouched
    r = O.valid_move_batch(s, segs, none, -1, 0.01, mover((2.5, 2.5, 0.1), (2.5, 2.5, 0.1)))[0]
    assert (r["x"], r["y"], r["z"], r["collided"]) == (2.5, 2.5, 0.1, 0)
    r = O.valid_move_batch(s, segs, none, -1, 0.01, mover((2.5, 2.5, 0.0), (9.5, 2.5, 0.0), radius=0.0))[0]
    assert (r["x"], r["collided"]) == (2.5, 0)


def test_valid_move_wall_and_floor_clamp():
    scene, segs = small_scene()
    s = O.Scene(scene)
    # through the bar at x = 5: west and east faces -> two intersection points, cannot move, nearest = west face
    r = O.valid_move_batch(s, segs, colliders(), -1, 0.01, mover((2.5, 4.5, 0.0), (8.5, 4.5, 0.0)))[0]
    assert (r["x"], r["y"], r["collided"], r["n_points"], r["n_collisions"]) == (2.5, 4.5, 1, 2, 0)
    assert (r["near_x"], r["near_y"]) == (5.0, 4.5)
    # going under the floor: a bottom-anchored entity is clamped to z = 0 and reports a collision (:346-352)
    r = O.valid_move_batch(s, segs, colliders(), -1, 0.01, mover((2.5, 2.5, 0.5), (2.75, 2.5, -0.25)))[0]
    assert (r["x"], r["z"], r["collided"]) == (2.75, 0.0, 1)
    # centre-anchored: zMin = |0 - h/2|
    r = O.valid_move_batch(s, segs, colliders(), -1, 0.01, mover((2.5, 2.5, 0.5), (2.75, 2.5, 0.1), height=0.5, anchor=1))[0]
    assert (r["z"], r["collided"]) == (0.25, 1)


def test_valid_move_sprite_circle_and_z_filter():
    scene, segs = small_scene()
    s = O.Scene(scene)
    c = colliders((3.5, 2.5, 0.0, 0.3, 1.0, 0, 0))
    # the move segment enters the combined circle (r = 0.5) at x = 3.0; from there toward the sprite centre the
    # sprite circle (r = 0.3) is met at x = 3.2
    r = O.valid_move_batch(s, segs, c, -1, 0.01, mover((2.5, 2.5, 0.0), (3.4, 2.5, 0.0)))[0]
    assert (r["x"], r["y"], r["collided"], r["n_collisions"], r["first_entity"]) == (2.5, 2.5, 1, 1, 0)
    assert abs(r["first_x"] - 3.2) < 1e-12 and abs(r["first_y"] - 2.5) < 1e-12
    assert r["first_z"] == 0.25                     # midZ of the mover clamped into the target's [0, 1]
    # same move, flying over it (z overlap fails): free
    r = O.valid_move_batch(s, segs, c, -1, 0.01, mover((2.5, 2.5, 2.0), (3.4, 2.5, 2.0)))[0]
    assert (r["x"], r["collided"], r["n_collisions"]) == (3.4, 0, 0)
    # starts and ends inside the combined circle without crossing it: "inside the circle" fallback (:178-189)
    r = O.valid_move_batch(s, segs, c, -1, 0.01, mover((3.3, 2.5, 0.0), (3.35, 2.5, 0.0)))[0]
    assert (r["collided"], r["n_collisions"], r["first_entity"], r["first_x"], r["first_y"]) == (1, 1, 0, 3.35, 2.5)
    # the entity's own collider and its parent's are skipped; destroyed / zero-radius colliders too
    r = O.valid_move_batch(s, segs, c, -1, 0.01, mover((2.5, 2.5, 0.0), (3.4, 2.5, 0.0), self_i=0))[0]
    assert r["collided"] == 0
    r = O.valid_move_batch(s, segs, c, -1, 0.01, mover((2.5, 2.5, 0.0), (3.4, 2.5, 0.0), parent=0))[0]
    assert r["collided"] == 0
    dead = colliders((3.5, 2.5, 0.0, 0.3, 1.0, 0, 1), (3.5, 2.5, 0.0, 0.0, 1.0, 0, 0))
    r = O.valid_move_batch(s, segs, dead, -1, 0.01, mover((2.5, 2.5, 0.0), (3.4, 2.5, 0.0)))[0]
    assert r["collided"] == 0


def test_valid_move_player_block_repeats_every_intersect_point():
    """collision.go:123-127 appends one player collision per element of intersectPoints — the wall points and
    the points of earlier iterations included.  Reproduced, not fixed."""
    scene, segs = small_scene()
    s = O.Scene(scene)
    c = colliders((3.5, 2.5, 0.0, 0.3, 1.0, 0, 0))     # the player
    # crosses the whole combined circle: 2 entry/exit points -> chk 1 adds 1 point (records: 1), chk 2 adds 1 (records: 2)
    r = O.valid_move_batch(s, segs, c, 0, 0.01, mover((2.5, 2.5, 0.0), (4.5, 2.5, 0.0)))[0]
    assert (r["n_points"], r["n_collisions"], r["first_entity"]) == (2, 3, 0)
    # a destroyed projectile, the player itself and the player's own projectiles skip the block (:100)
    for kw in (dict(destroyed=1), dict(self_i=0), dict(parent=0)):
        r = O.valid_move_batch(s, segs, c, 0, 0.01, mover((2.5, 2.5, 0.0), (4.5, 2.5, 0.0), **kw))[0]
        assert r["n_collisions"] == 0 and r["collided"] == 0


def test_check_alternate_slides_along_a_wall():
    scene, segs = small_scene()
    s = O.Scene(scene)
    # heading east into the bar's west face with a little northward drift: X-only fails, Y-only succeeds
    m = mover((4.9, 4.5, 0.0), (5.1, 4.6, 0.0))
    r = O.valid_move_alternate(s, segs, colliders(), -1, 0.01, m, heading=0.0, velocity=0.2)
    assert r["collided"] == 1 and (r["x"], r["y"]) == (4.9, 4.6)
    # without checkAlternate it just stays
    r = O.valid_move_alternate(s, segs, colliders(), -1, 0.01, m, heading=0.0, velocity=0.2, check_alternate=False)
    assert (r["x"], r["y"], r["collided"]) == (4.9, 4.5, 1)


def test_host_helpers():
    assert col.angle_distance(0.0, 0.5) == math.fmod(math.fmod(0.5 + math.pi, 2 * math.pi) - 2 * math.pi, 2 * math.pi) + math.pi
    assert abs(col.angle_distance(0.1, 0.3) - 0.2) < 1e-12 and abs(col.angle_distance(0.3, 0.1) + 0.2) < 1e-12
    assert col.nearly_equal(1.0, 1.0 + 1e-4, 0.001) and not col.nearly_equal(1.0, 1.01, 0.001)
    assert col.nearly_equal(0.0, 0.0, 0.001) and not col.nearly_equal(0.0, 1e-9, 0.001)
