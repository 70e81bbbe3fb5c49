"""CPU tests of the Pathing.FindPath oracle (oracle/pm3d_oracle_path.c): hand-derived answers of the parts the
reference pins ([V]: neighbour order, blockers, the segmentation into curves, bezier.New / Curve.Point) and the
invariants any A* result must satisfy.  The heap tie-breaking is [R] (go-astar + container/heap are not on disk)."""
import numpy as np

from oracle import oracle as O
from pm3d import workloads as W


def room():
    lvl = np.zeros((12, 10), dtype=np.uint16)
    lvl[0, :] = lvl[-1, :] = 1
    lvl[:, 0] = lvl[:, -1] = 1
    lvl[5, 1:7] = 1
    return lvl


def test_short_paths_are_returned_as_is():
    lvl = room()
    pts, cnt, cells, _ = O.find_path_batch(lvl, [[1.5, 1.5, 1.9, 1.1], [1.2, 1.7, 2.5, 1.5]], cap=16, threads=1)
    assert cnt.tolist() == [1, 2] and cells.tolist() == [1, 2]        # fewer than 3 steps: cell centres, no curve
    assert pts[0, 0].tolist() == [1.5, 1.5]
    assert pts[1, :2].tolist() == [[1.5, 1.5], [2.5, 1.5]]


def test_three_collinear_steps_give_the_bezier_of_three_control_points():
    pts, cnt, cells, _ = O.find_path_batch(room(), [[1.5, 1.5, 3.5, 1.5]], cap=16, threads=1)
    assert cells[0] == 3 and cnt[0] == 6                               # 2 * segmentSize points
    # quadratic Bezier through (1.5,1.5) (2.5,1.5) (3.5,1.5) at t = k/5, evaluated the reference's way
    def curve(c3):
        out = []
        for k in range(6):
            t = k / 5.0
            c = [c3[0] * 1.0, c3[1] * 2.0, c3[2] * 1.0]                    # control = point * binomial
            p0, p1, p2 = c[0], c[1] * t, c[2] * (t * t)
            t1 = 1 - t
            x = p2
            x += p1 * t1
            x += p0 * (t1 * t1)
            out.append(x)
        return out
    assert pts[0, :6, 0].tolist() == curve([1.5, 2.5, 3.5])
    assert pts[0, :6, 1].tolist() == curve([1.5, 1.5, 1.5])
    assert np.allclose(pts[0, :6, 1], 1.5, rtol=0, atol=1e-15)


def test_no_path_and_outside():
    lvl = room()
    lvl[1:5, 5] = 1; lvl[1, 5] = 1
    lvl[6:11, 7] = 1
    closed = lvl.copy(); closed[2, 1:9] = 1                            # wall across the room
    pts, cnt, _, exp = O.find_path_batch(closed, [[1.5, 1.5, 8.5, 5.5], [-3.0, 2.0, 4.0, 4.0], [1.5, 1.5, 40.0, 4.0]], cap=8, threads=1)
    assert cnt.tolist() == [-1, -1, -1]
    assert exp[0] > 0 and exp[1] == 0                                  # the first one searched its whole component


def test_path_invariants_on_the_bundled_map():
    g = W.load_golden("proving_grounds")
    lvl = g["levels"][0]
    q = W.path_queries(lvl, 48, 5)
    pts, cnt, cells, exp = O.find_path_batch(lvl, q, cap=8192, threads=8)
    assert (cnt > 0).all()
    for i in range(len(q)):
        p = pts[i, :cnt[i]]
        assert tuple(p[0]) == (int(q[i, 0]) + 0.5, int(q[i, 1]) + 0.5)   # curves start and end on the end cells' centres
        assert tuple(p[-1]) == (int(q[i, 2]) + 0.5, int(q[i, 3]) + 0.5)
        manhattan = abs(int(q[i, 0]) - int(q[i, 2])) + abs(int(q[i, 1]) - int(q[i, 3]))
        assert cells[i] >= manhattan + 1                                 # a 4-connected path cannot be shorter
        assert exp[i] >= cells[i]
    # the open map: most paths are as short as the Manhattan distance allows
    short = sum(cells[i] == abs(int(q[i, 0]) - int(q[i, 2])) + abs(int(q[i, 1]) - int(q[i, 3])) + 1 for i in range(len(q)))
    assert short > len(q) // 2


def test_segmentation_lengths():
    """len(path) = L gives 2 * sum(segment sizes) points: segments of 7 sharing their end steps, a remainder of 1-2
    steps pulled into the last curve (pathing.go:157-181)."""
    lvl = np.zeros((40, 3), dtype=np.uint16)
    for L in range(3, 31):
        pts, cnt, cells, _ = O.find_path_batch(lvl, [[0.5, 1.5, L - 0.5, 1.5]], cap=256, threads=1)
        assert cells[0] == L
        seg, idx, size, rem = [], 0, 7, L
        while rem > 0:
            if size <= 1:
                seg.append(0.5); rem -= 1; idx += 1
                continue
            if rem < size:
                size = rem
            rem -= size - 1
            if 0 < rem < 3:
                size += rem - 1; rem = 0
            seg.append(size)
            idx += size - 1
        assert cnt[0] == int(2 * sum(seg)), (L, seg, cnt[0])
