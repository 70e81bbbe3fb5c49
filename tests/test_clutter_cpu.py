"""CPU tests of the clutter oracle (oracle/pm3d_oracle_clutter.c): Go's math/rand restated and PINNED against the
outputs of the real library that every Go programmer has seen, ClutterHandler.Update against hand-derived answers,
and the host-side pieces of pm3d/clutter.py that need no GPU."""
import math

import numpy as np
import pytest

import pm3d
from oracle import oracle as O
from pm3d import clutter as cl


# ------------------------------------------------------------------ Go math/rand: known answers of the real library

def test_go_rand_seed_1_known_answers():
    """rand.Seed(1) (the default source before Go 1.20): the sequences printed by the Go tour, the package examples
    and countless bug reports ("why is my random number always 81?")."""
    r = O.GoRand(1)
    assert [r.intn(100) for _ in range(10)] == [81, 87, 47, 59, 81, 18, 25, 40, 56, 0]
    r = O.GoRand(1)
    assert r.intn(10) == 1                                   # "My favorite number is 1" (A Tour of Go)
    r = O.GoRand(1)
    assert [r.int63() for _ in range(2)] == [5577006791947779410, 8674665223082153551]
    r = O.GoRand(1)
    assert [r.float64() for _ in range(5)] == [0.6046602879796196, 0.9405090880450124, 0.6645600532184904,
                                                0.4377141871869802, 0.4246374970712657]


def test_go_rand_cooked_table():
    ck = O.gorand_cooked()
    assert ck.view(np.int64)[0] == -4181792142133755926      # first two entries of rngCooked in math/rand/rng.go
    assert ck.view(np.int64)[1] == -4576982950128230565
    # the jump-ahead equals stepping the generator one addition at a time
    for n in (0, 1, 333, 334, 606, 607, 608, 5000, 123457):
        assert np.array_equal(O.gorand_gen_state(n, False), O.gorand_gen_state(n, True)), n


def test_library_regenerates_the_same_table():
    """libpm3d builds rngCooked with its own host code (csrc/clutter_kernel.cu); no GPU involved."""
    lib = pm3d.load()
    out = np.zeros(607, dtype=np.uint64)
    assert lib.pm3d_go_rand_cooked(out.ctypes.data) == 0
    assert np.array_equal(out, O.gorand_cooked())


def test_go_rand_seed_normalisation():
    """Seed: seed % (2^31-1), negative += 2^31-1, zero -> 89482311."""
    m = 2147483647
    def first(seed):
        return O.GoRand(seed).int63()
    assert first(1) == first(1 + m) == first(1 + 5 * m) == first(1 - m)
    assert first(0) == first(m) == first(-m) == first(89482311)
    assert first(-1) == first(m - 1)
    assert first(2) != first(1)


def test_draw_k_depends_on_two_initial_words_only():
    """What the device kernel relies on: for k <= 273 the k-th output of a fresh source is vec[334-k] + vec[607-k] of
    the initial state."""
    r = O.GoRand(987654321)
    init = np.array(r._s.vec[:], dtype=np.uint64)
    for k in range(1, 274):
        want = (int(init[334 - k]) + int(init[607 - k])) & ((1 << 64) - 1)
        assert r.uint64() == want, k
    # ... and no further: draw 274 reads a word that draw 1 overwrote
    want = (int(init[334 - 274]) + int(init[607 - 274])) & ((1 << 64) - 1)
    assert r.uint64() != want


# ------------------------------------------------------------------ ClutterHandler.Update, hand-derived

def tiny_map():
    lvl = np.zeros((8, 6), dtype=np.uint16)
    lvl[0, :] = lvl[-1, :] = 1
    lvl[:, 0] = lvl[:, -1] = 1
    lvl[4, 2] = 3
    cls = np.zeros((8, 6), dtype=np.uint8)
    cls[3:, :] = 1                                           # x >= 3: second floor path
    return lvl, cls


def defs_of(*rows):
    d = np.zeros(len(rows), dtype=O.CLUTTER_DEF_DTYPE)
    for i, (freq, scale) in enumerate(rows):
        d[i] = (freq, scale, i, 0, 0, 17, 17, 0)
    return d


def query(x, y, heading, fov):
    ang = O.clutter_angles(heading, fov)
    q = np.zeros(1, dtype=O.CLUTTER_QUERY_DTYPE)
    q["cam_x"], q["cam_y"], q["n_angles"] = x, y, len(ang)
    q["cs"][0, :len(ang), 0] = np.cos(ang)
    q["cs"][0, :len(ang), 1] = np.sin(ang)
    return q


def test_update_visits_the_cells_along_the_fan():
    lvl, cls = tiny_map()
    masks = np.array([0b1, 0b1], dtype=np.uint64)
    # looking along +x from (1.5, 2.5) with a very narrow fan: cells (1,2) (2,2) (3,2); (4,2) is a wall, (5,2) (6,2)
    # lie behind it but the reference does not test visibility; (7,2) is the boundary wall
    got = O.clutter_update(lvl, cls, masks, defs_of((1.0, 0.05)), 7, 6.0, query(1.5, 2.5, 0.0, 1e-6))
    assert [(int(r["cell_x"]), int(r["cell_y"])) for r in got] == [(1, 2), (2, 2), (3, 2), (5, 2), (6, 2)]
    for r in got:                                            # float64(x) + Float64() stays inside the cell
        assert r["cell_x"] <= r["x"] < r["cell_x"] + 1 and r["cell_y"] <= r["y"] < r["cell_y"] + 1 and r["def"] == 0
    # the positions are the 2nd and 3rd draw of rand seeded with Map.Seed + posId, posId = (x-1)*mapWidth + y
    for r in got:
        g = O.GoRand(7 + (int(r["cell_x"]) - 1) * 8 + int(r["cell_y"]))
        assert g.float64() <= 1.0
        assert (r["x"], r["y"]) == (r["cell_x"] + g.float64(), r["cell_y"] + g.float64())


def test_update_frequency_mask_and_early_returns():
    lvl, cls = tiny_map()
    q = query(2.5, 2.5, 0.3, 70.0)
    both = np.array([0b11, 0b11], dtype=np.uint64)
    none = O.clutter_update(lvl, cls, both, defs_of((-1.0, 1.0), (-1.0, 1.0)), 1, 4.0, q)
    assert len(none) == 0                                    # Float64() <= -1 never holds
    every = O.clutter_update(lvl, cls, both, defs_of((1.0, 1.0), (1.0, 1.0)), 1, 4.0, q)
    cells = {(int(r["cell_x"]), int(r["cell_y"])) for r in every}
    assert len(every) == 2 * len(cells)                      # both clutter types in every visited cell
    assert all(lvl[c] == 0 for c in cells)
    # FloorPathMatch: clutter 1 only on the second floor path (x >= 3)
    split = O.clutter_update(lvl, cls, np.array([0b01, 0b11], dtype=np.uint64), defs_of((1.0, 1.0), (1.0, 1.0)), 1, 4.0, q)
    assert all(r["cell_x"] >= 3 for r in split if r["def"] == 1) and any(r["def"] == 1 for r in split)
    assert {(int(r["cell_x"]), int(r["cell_y"])) for r in split if r["def"] == 0} == cells
    # a skipped clutter entry draws nothing: clutter 1's numbers on x < 3 cells move up when clutter 0 is masked out
    only1 = O.clutter_update(lvl, cls, np.array([0b10, 0b10], dtype=np.uint64), defs_of((1.0, 1.0), (1.0, 1.0)), 1, 4.0, q)
    first0 = {(int(r["cell_x"]), int(r["cell_y"])): (r["x"], r["y"]) for r in every if r["def"] == 0}
    assert all(first0[(int(r["cell_x"]), int(r["cell_y"]))] == (r["x"], r["y"]) for r in only1)
    # clutter.go:31-34
    assert len(O.clutter_update(lvl, cls, both, defs_of((1.0, 1.0)), 1, 0.0, q)) == 0
    assert len(O.clutter_update(lvl, cls, both, defs_of(), 1, 4.0, q)) == 0


def test_a_cell_gets_the_same_clutter_from_every_camera():
    lvl, cls = tiny_map()
    masks = np.array([0b11, 0b11], dtype=np.uint64)
    d = defs_of((0.6, 1.0), (0.4, 1.0))
    a = O.clutter_update(lvl, cls, masks, d, 99, 5.0, query(1.5, 1.5, 0.7, 70.0))
    b = O.clutter_update(lvl, cls, masks, d, 99, 5.0, query(5.5, 4.2, 3.9, 100.0))
    da, db = {}, {}
    for recs, dst in ((a, da), (b, db)):
        for r in recs:
            dst.setdefault((int(r["cell_x"]), int(r["cell_y"])), []).append((int(r["def"]), float(r["x"]), float(r["y"])))
    common = set(da) & set(db)
    assert len(common) >= 3
    assert all(da[c] == db[c] for c in common)


def test_camera_outside_the_map_and_negative_coordinates():
    lvl, cls = tiny_map()
    lvl[:] = 0                                               # no walls: int64(-0.5) = 0 lands in column 0
    masks = np.array([1, 1], dtype=np.uint64)
    got = O.clutter_update(lvl, cls, masks, defs_of((1.0, 1.0)), 3, 3.0, query(-0.5, 2.5, 0.0, 1e-6))
    assert [(int(r["cell_x"]), int(r["cell_y"])) for r in got] == [(0, 2), (1, 2), (2, 2)]   # -0.5 and 0.5 share cell 0
    got = O.clutter_update(lvl, cls, masks, defs_of((1.0, 1.0)), 3, 3.0, query(-1.5, 2.5, 0.0, 1e-6))
    assert [(int(r["cell_x"]), int(r["cell_y"])) for r in got] == [(0, 2), (1, 2)]          # -1.5 is outside (x < 0)


# ------------------------------------------------------------------ host mirror pieces

@pytest.mark.parametrize("fov", [20.0, 45.0, 60.0, 70.0, 90.0, 110.0, 179.0])
def test_sample_angles_match_the_oracle_loop(fov):
    counts = set()
    for k in range(400):
        heading = (k * 0.6180339887498949 * 2 * math.pi) % (2 * math.pi)
        a = cl.sample_angles(heading, fov)
        b = O.clutter_angles(heading, fov)
        assert a == list(b)
        counts.add(len(a))
    assert counts <= {20, 21} and 21 in counts               # the float accumulation decides between 20 and 21


def test_resolve_clutter_masks():
    clutter = [cl.MapClutter("rocks/rock_0.png", 0.3, 1.0, "grass"), cl.MapClutter("shrubbery/bush_0.png", 0.5, 1.0, "grass"),
               cl.MapClutter("x.png", 0.1, 2.0, None), cl.MapClutter("y.png", 0.1, 2.0, r"floor_\d")]
    sheets = {"rocks/rock_0.png": (4, 18, 18), "shrubbery/bush_0.png": (5, 24, 24), "x.png": (6, 8, 9), "y.png": (7, 3, 3)}
    defs, cls, masks = cl.resolve_clutter(clutter, np.zeros((4, 4)), ["floors/grass.png", "floors/floor_1a.png"], sheets)
    assert masks.tolist() == [0b0111, 0b1100]
    assert defs["scale"].tolist() == [0.05, 0.05, 0.1, 0.1]
    assert (defs["rx1"].tolist(), defs["ry1"].tolist()) == ([17, 23, 7, 2], [17, 23, 8, 2])   # 1 px short (sprite.go:176)
    assert cls.dtype == np.uint8
