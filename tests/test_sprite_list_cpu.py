"""CPU tests of the getRaycastSprites restatement (oracle/pm3d_oracle_clutter.c) against hand-derived answers."""
import numpy as np

from oracle import oracle as O


def sprites_at(*xy):
    s = np.zeros(len(xy), dtype=O.SPRITE_DTYPE)
    for i, (x, y) in enumerate(xy):
        s[i]["x"], s[i]["y"], s[i]["scale"], s[i]["sheet"] = x, y, 1.0, i
    return s


def test_point_in_proximity_uses_the_ceiling_of_the_distance():
    # renderDistance 2.3 -> ceil = 3: |dx| <= 3 and |dy| <= 3 (game/model/math.go:93-96)
    world = sprites_at((13.0, 10.0), (13.0000001, 10.0), (7.0, 7.0), (10.0, 13.5), (10.0, 10.0), (6.9999, 12.0))
    got, idx = O.raycast_sprites(world, 10.0, 10.0, 2.3)
    assert idx.tolist() == [0, 2, 4]
    assert np.array_equal(got, world[[0, 2, 4]])
    # an integral distance is its own ceiling
    assert O.raycast_sprites(world, 10.0, 10.0, 3.0)[1].tolist() == [0, 2, 4]
    assert O.raycast_sprites(world, 10.0, 10.0, 3.5)[1].tolist() == [0, 1, 2, 3, 4, 5]


def test_negative_distance_target_clutter_and_extras():
    world = sprites_at((50.0, 50.0), (1.0, 1.0), (99.0, 3.0))
    assert O.raycast_sprites(world, 0.0, 0.0, -1.0)[1].tolist() == [0, 1, 2]       # renderDistance < 0: everything
    assert O.raycast_sprites(world, 0.0, 0.0, 5.0)[1].tolist() == [1]
    assert O.raycast_sprites(world, 0.0, 0.0, 5.0, target=2)[1].tolist() == [1, 2]  # the player's target is always listed
    clutter, extras = sprites_at((0.2, 0.3), (0.4, 0.5)), sprites_at((7.0, 7.0))
    got, idx = O.raycast_sprites(world, 0.0, 0.0, 5.0, -1, clutter, extras)
    assert idx.tolist() == [1, -1, -2, -2147483648]                                 # map sprites, clutter, extras
    assert np.array_equal(got[1:3], clutter) and np.array_equal(got[3:], extras)
    assert len(O.raycast_sprites(world[:0], 0.0, 0.0, 5.0)[0]) == 0
    nan = sprites_at((float("nan"), 1.0))
    assert len(O.raycast_sprites(nan, 0.0, 0.0, 5.0)[0]) == 0                       # Go: comparisons with NaN are false
