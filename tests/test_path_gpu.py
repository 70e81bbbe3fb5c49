"""GPU parity of pm3d_find_path_batch against the oracle's restatement of Pathing.FindPath
(game/model/pathing.go:123-186 over go-astar + container/heap): the same path cells — the heap's tie-breaking
included — and bit-identical curve points."""
import numpy as np
import pytest

import pm3d
from oracle import oracle as O
from pm3d import workloads as W
from pm3d.pathing import Pathing

pytestmark = pytest.mark.gpu


def check(ctx, lvl, q, cap=8192):
    grid = np.ascontiguousarray(lvl, dtype=np.uint16)[None]
    ctx.upload_scene(W.scene_from_grid(grid, 1))
    got_p, got_n = ctx.find_path_batch(q, cap)
    want_p, want_n, _, _ = O.find_path_batch(lvl, q, cap=cap, threads=8)
    assert np.array_equal(got_n, want_n)
    for i in range(len(q)):
        k = max(int(want_n[i]), 0)
        assert got_p[i, :k].tobytes() == want_p[i, :k].tobytes(), f"query {i}: {q[i].tolist()}"
    return want_n


@pytest.mark.parametrize("name", ["proving_grounds", "debug_land"])
def test_bundled_maps(gpu_ctx, name):
    lvl = W.load_golden(name)["levels"][0]
    n = check(gpu_ctx, lvl, W.path_queries(lvl, 96, 11))
    assert (n > 0).all()


@pytest.mark.parametrize("density,seed", [(0.10, 42), (0.30, 7), (0.45, 8)])
def test_synthetic_mazes(gpu_ctx, density, seed):
    """Dense random maps: detours, dead ends, unreachable targets (count -1), re-opened queue entries."""
    lvl = W.synthetic_levels(96, 80, 1, density, seed)[0]
    q = W.path_queries(lvl, 256, seed + 1)
    q[0, 2:] = q[0, :2]                       # start == finish
    q[1] = (-2.0, 3.0, 5.5, 5.5)              # nil start tile
    q[2] = (5.5, 5.5, 500.0, 5.5)             # nil finish tile
    walls = np.argwhere(lvl != 0)
    q[3, 2:] = walls[len(walls) // 2] + 0.5   # finish inside a wall: never reached
    q[4, :2] = walls[len(walls) // 3] + 0.5   # start inside a wall: allowed (the start tile is never tested)
    n = check(gpu_ctx, lvl, q)
    assert (n > 0).any() and (n[1:4] == -1).all()
    if density >= 0.45:
        assert (n == -1).sum() > 8            # disconnected components occur


def test_output_capacity_is_reported(gpu_ctx):
    lvl = np.zeros((64, 3), dtype=np.uint16)
    grid = lvl[None]
    gpu_ctx.upload_scene(W.scene_from_grid(grid, 1))
    q = np.array([[0.5, 1.5, 60.5, 1.5]])
    pts, n = gpu_ctx.find_path_batch(q, 16)
    want_p, want_n, _, _ = O.find_path_batch(lvl, q, cap=512, threads=1)
    assert n[0] == want_n[0] > 16                                  # the full count, although only 16 points fit
    assert pts[0].tobytes() == want_p[0, :16].tobytes()
    full, err = Pathing(gpu_ctx, max_points=16).FindPath((0.5, 1.5), (60.5, 1.5))   # the mirror asks again, larger
    assert err is None and full.tobytes() == want_p[0, :want_n[0]].tobytes()
    _, err = Pathing(gpu_ctx).FindPath((0.5, 1.5), (90.0, 1.5))
    assert err == "unable to find path for (0,2) -> (90,2)"


def test_many_queries_in_waves(gpu_ctx):
    """More queries than one wave of searches holds on a 500x500 map (16 bytes of scratch per cell and search)."""
    lvl = W.load_golden("proving_grounds")["levels"][0]
    q = W.path_queries(lvl, 1500, 3, max_manhattan=40)
    check(gpu_ctx, lvl, q, cap=512)
