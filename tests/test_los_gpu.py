"""GPU parity of the batched line-of-sight / wall-collision queries against the oracle's
brute-force restatement of game/collision.go:40-65 and :69-97,202-203,307-342.  Booleans and
positions are exact (bit-for-bit): the device runs the same binary64 predicate, only on fewer
candidate segments."""
import numpy as np
import pytest

import pm3d
from oracle import oracle as O
from pm3d import workloads as W

pytestmark = pytest.mark.gpu


def c5_small(n_rays, seed=42, size=256, density=0.10):
    grid = W.synthetic_levels(size, size, 1, density, seed)
    segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
    rays = W.los_rays(grid[0], n_rays, seed + 1)
    return grid, segs, rays


def test_c5_map_los_parity(gpu_ctx):
    grid, segs, rays = c5_small(20000)
    gpu_ctx.upload_segments(segs)
    got = gpu_ctx.los_batch(rays)
    want = O.los_batch(segs, rays, threads=8)
    assert np.array_equal(got, want)
    assert 0 < int(want.sum()) < len(want)  # both outcomes occur


@pytest.mark.parametrize("name", ["proving_grounds", "debug_land"])
def test_bundled_map_los_parity(gpu_ctx, name):
    g = W.load_golden(name)
    segs = g["segments"]
    gpu_ctx.upload_segments(segs)
    lvl = g["levels"][0]
    rays = W.los_rays(lvl[:64, :64], 30000, 3)
    got = gpu_ctx.los_batch(rays)
    assert np.array_equal(got, O.los_batch(segs, rays, threads=8))
    # long rays across the whole map
    rays = W.los_rays(lvl, 20000, 4)
    assert np.array_equal(gpu_ctx.los_batch(rays), O.los_batch(segs, rays, threads=8))


def test_los_edge_rays(gpu_ctx):
    """Zero-length rays, rays ending exactly on wall faces / corners / cell borders, axis-aligned rays
    running along a wall face, rays leaving the map and non-finite input."""
    lvl = np.zeros((12, 10), dtype=np.uint16)
    lvl[0, :] = lvl[-1, :] = 32767
    lvl[:, 0] = lvl[:, -1] = 32767
    lvl[5, 3:7] = 1
    lvl[7, 5] = 2
    segs = pm3d.generate_wall_collision_lines(lvl, True, W.CLIP_DISTANCE)
    rays = [
        [2.5, 4.5, 2.5, 4.5],          # zero length
        [2.5, 4.5, 5.0, 4.5],          # ends exactly on a west face
        [2.5, 3.0, 8.5, 3.0],          # runs along the bar's south face (collinear: denom == 0 there)
        [2.5, 2.0, 5.0, 3.0],          # ends exactly on a wall corner
        [5.0, 1.5, 5.0, 8.5],          # runs along the bar's west face line x = 5
        [6.0, 7.5, 6.0, 2.5],          # along its east face
        [2.5, 4.5, 9.5, 4.5],          # through the bar
        [1.5, 1.5, 10.5, 8.5],         # long diagonal
        [1.005, 1.005, 10.995, 8.995],  # inside the inner border rect inset by clip
        [1.5, 1.5, 1.5, 0.5],          # crosses the inner border rect into the boundary ring
        [-4.0, 5.0, 20.0, 5.0],        # starts and ends outside the map
        [3.5, 3.5, 3.5, 1e9],          # absurdly long
        [2.0, 2.0, 4.0, 4.0],          # passes through lattice corner (3,3) exactly
        [4.0, 2.0, 6.0, 8.0],
        [float("nan"), 2.0, 4.0, 4.0],
        [2.0, 2.0, float("inf"), 4.0],
        [8.999999999999998, 4.5, 2.5, 4.5000000000000009],
    ]
    rays = np.array(rays, dtype=np.float64)
    gpu_ctx.upload_segments(segs)
    got = gpu_ctx.los_batch(rays)
    want = O.los_batch(segs, rays)
    assert got.tolist() == want.tolist()


def test_los_grid_aligned_stress(gpu_ctx):
    """Endpoints on the half/quarter/integer lattice: many rays pass exactly through cell corners and
    along wall faces, where the cell walk and the exact predicate are most fragile."""
    grid = W.synthetic_levels(48, 48, 1, 0.25, 11)
    segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
    rng = W.SplitMix64(12)
    u = rng.uniform(4 * 40000).reshape(-1, 4)
    rays = np.round(u * 47 * 4) / 4.0 + 0.5
    gpu_ctx.upload_segments(segs)
    got = gpu_ctx.los_batch(rays)
    want = O.los_batch(segs, rays, threads=8)
    assert np.array_equal(got, want)


def test_los_no_segments_and_empty_batch(gpu_ctx):
    gpu_ctx.upload_segments(np.zeros((0, 4)))
    rays = np.array([[0.5, 0.5, 3.5, 3.5]])
    assert gpu_ctx.los_batch(rays).tolist() == [1]
    assert len(gpu_ctx.los_batch(np.zeros((0, 4)))) == 0


def test_move_batch_parity(gpu_ctx):
    grid, segs, rays = c5_small(20000, seed=7, size=96, density=0.15)
    scene = W.scene_from_grid(grid, 1)
    gpu_ctx.upload_scene(scene)
    gpu_ctx.upload_segments(segs)
    # short moves (a few tenths of a cell), some no-ops, some leaving the map
    rng = W.SplitMix64(8)
    d = (rng.uniform(2 * len(rays)).reshape(-1, 2) - 0.5) * 1.6
    moves = rays.copy()
    moves[:, 2:] = moves[:, :2] + d
    moves[::97, 2:] = moves[::97, :2]            # no-op moves
    moves[5, 2:] = (-3.0, 4.0)
    moves[6, 2:] = (500.0, 4.0)
    moves[7, :2] = (1.5, 1.5); moves[7, 2:] = (1.5, -7.0)
    pos, col = gpu_ctx.move_batch(moves, W.CLIP_DISTANCE)
    wpos, wcol, _ = O.move_walls_batch(O.Scene(scene), segs, moves, W.CLIP_DISTANCE, threads=8)
    assert np.array_equal(col, wcol)
    assert np.array_equal(pos.view(np.uint64), wpos.view(np.uint64))
    assert 0 < int(wcol.sum()) < len(wcol)


def test_full_size_16m_rays_properties(gpu_ctx):
    """BASELINE config C5 at its full size: 16 Mi rays.  The oracle checks a 1/256 sample; the whole
    batch is checked through size-independent properties: determinism, batch-split invariance and
    the rate of blocked rays of the sample matching the full batch."""
    grid, segs, _ = c5_small(16)
    n = 1 << 24
    rays = W.los_rays(grid[0], n, 43)
    gpu_ctx.upload_segments(segs)
    full = gpu_ctx.los_batch(rays)
    sample = rays[::256]
    assert np.array_equal(full[::256], O.los_batch(segs, sample, threads=8))
    # split invariance: two halves in separate calls
    half = n // 2
    assert np.array_equal(gpu_ctx.los_batch(rays[:half]), full[:half])
    assert np.array_equal(gpu_ctx.los_batch(rays[half:]), full[half:])
    assert abs(float(full.mean()) - float(full[::256].mean())) < 0.01


def test_division_free_predicate_matches_reference_arithmetic(gpu_ctx):
    """The device decides t,u in [0,1] from signs and magnitudes instead of dividing (los_kernel.cu
    segments_intersect).  Adversarial inputs: endpoints exactly on segments (t or u == 0 or 1), collinear and
    parallel rays, one-ulp perturbations around touching configurations, huge / tiny / zero / non-finite
    coordinates (where it must fall back to the reference arithmetic, including Go's NaN comparisons)."""
    segs = np.array([
        [2.0, 1.0, 2.0, 5.0],          # vertical
        [1.0, 3.0, 6.0, 3.0],          # horizontal
        [0.5, 0.5, 4.5, 6.5],          # diagonal
        [3.0, 3.0, 3.0, 3.0],          # degenerate point segment
        [7.25, 0.125, 7.25, 0.375],    # tiny
    ])
    rng = W.SplitMix64(2024)
    rays = []
    for sg in segs:
        x3, y3, x4, y4 = sg
        for a in (0.0, 1.0, 0.5, 0.25, 1.0 / 3.0):
            px, py = x3 + a * (x4 - x3), y3 + a * (y4 - y3)    # a point on the segment
            for dxy in ((1.0, 0.0), (0.0, 1.0), (1.0, 1.0), (-2.5, 0.75), (1e-9, 1.0)):
                for k in (1.0, 1e-3, 1e3):
                    ex, ey = px + k * dxy[0], py + k * dxy[1]
                    rays += [[px, py, ex, ey], [ex, ey, px, py]]                     # starts / ends exactly on it
                    rays += [[np.nextafter(px, 9), py, ex, ey], [np.nextafter(px, -9), py, ex, ey],
                             [px, np.nextafter(py, 9), ex, ey], [ex, ey, px, np.nextafter(py, -9)]]
                    rays += [[2 * px - ex, 2 * py - ey, ex, ey]]                     # passes through it
        rays += [[x3, y3, x4, y4], [x4, y4, x3, y3], [x3 - 1, y3 - 1, x3 - 1, y3 - 1]]  # collinear / identical / point
    u = rng.uniform(4 * 4000).reshape(-1, 4)
    rays += (u * 8.0).tolist()
    rays += (np.round(u[:1000] * 16) / 2.0).tolist()                              # lattice points: many exact touches
    big = [1e200, -1e200, 1e-200, -1e-200, 1e308, 5e-324, 0.0, -0.0, float("inf"), -float("inf"), float("nan")]
    for i, bx in enumerate(big):
        for by in big[i::3]:
            rays += [[bx, by, 3.0, 3.0], [3.0, 2.0, bx, by], [bx, 2.0, by, 4.0], [2.0, bx, 2.0, by]]
    rays = np.array(rays, dtype=np.float64)
    gpu_ctx.upload_segments(segs)
    got = gpu_ctx.los_batch(rays)
    with np.errstate(all="ignore"):
        want = O.los_batch(segs, rays)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first: {rays[bad[:5]].tolist()} gpu {got[bad[:5]].tolist()} oracle {want[bad[:5]].tolist()}"
    assert 0 < int(want.sum()) < len(want)


# ---- two-pass line of sight (k_los_fast + k_los): the first pass decides rays from per-cell code bytes alone —
#      "blocked" when the ray provably crosses the interior of a unit edge a segment covers, "visible" when no
#      visited cell has a candidate — and must give the brute-force booleans like the exact pass it replaces.
def _both_modes(ctx, rays, monkeypatch):
    """(first pass in its default mode, first pass with the verify mode forced the other way ... , exact pass alone):
    all modes of the first pass must agree with the exact pass, so the two are folded into one array pair."""
    monkeypatch.setenv("PM3D_LOS_FAST", "1")
    fast = ctx.los_batch(rays)
    for v in ("0", "1"):                      # verify mode off / on, whatever the upload chose
        monkeypatch.setenv("PM3D_LOS_VERIFY", v)
        forced = ctx.los_batch(rays)
        assert np.array_equal(forced, fast), f"PM3D_LOS_VERIFY={v} changes the answers"
    monkeypatch.delenv("PM3D_LOS_VERIFY")
    monkeypatch.setenv("PM3D_LOS_FAST", "0")
    exact = ctx.los_batch(rays)
    monkeypatch.delenv("PM3D_LOS_FAST")
    return fast, exact


@pytest.mark.parametrize("size,density,seed", [(256, 0.10, 42), (96, 0.30, 5), (200, 0.01, 9), (300, 0.002, 10), (64, 0.0, 3)])
def test_two_pass_los_matches_brute_force(gpu_ctx, monkeypatch, size, density, seed):
    """Dense maps (most rays blocked within a few cells), sparse maps (most rays visible, long walks through
    candidate-free cells) and an empty map (only the border rectangles)."""
    grid, segs, rays = c5_small(30000, seed=seed, size=size, density=density)
    gpu_ctx.upload_segments(segs)
    fast, exact = _both_modes(gpu_ctx, rays, monkeypatch)
    want = O.los_batch(segs, rays, threads=8)
    assert np.array_equal(exact, want)
    assert np.array_equal(fast, want)
    if density <= 0.01:
        assert want.mean() > 0.3  # the visible path is exercised


def test_two_pass_los_corner_grazes(gpu_ctx, monkeypatch):
    """Rays aimed at wall-cell corners and at points of wall faces within 2^-18 .. 2^-24 of a corner (either side of
    the first pass's 2^-20 margin), rays through lattice corners exactly, rays starting or ending ON a face, rays
    running along faces: the first pass must hand every one it cannot prove to the exact predicate."""
    grid = W.synthetic_levels(40, 40, 1, 0.2, 21)
    lvl = grid[0]
    segs = pm3d.generate_wall_collision_lines(lvl, True, W.CLIP_DISTANCE)
    walls = np.argwhere(lvl[1:-1, 1:-1] != 0) + 1
    free = np.argwhere(lvl == 0)
    rng = W.SplitMix64(77)
    rays = []
    offs = [0.0] + [s * 2.0 ** -e for e in (18, 19, 20, 21, 22, 24, 30, 40, 52) for s in (1.0, -1.0)]
    for wx, wy in walls[:160]:
        for cx, cy in ((wx, wy), (wx + 1, wy), (wx, wy + 1), (wx + 1, wy + 1)):
            src = free[(rng.uniform(3) * len(free)).astype(np.int64)] + 0.5
            for sx, sy in src:
                for o in offs:
                    # through / short of / beyond the point (corner + o along either axis)
                    for tx, ty in ((cx + o, cy), (cx, cy + o)):
                        rays.append([sx, sy, tx, ty])
                        rays.append([sx, sy, 2 * tx - sx, 2 * ty - sy])
                        rays.append([tx, ty, sx, sy])
    rays = np.array(rays, dtype=np.float64)
    rays = rays[(rays.min(axis=1) > 0.5) & (rays.max(axis=1) < 39.5)]
    gpu_ctx.upload_segments(segs)
    fast, exact = _both_modes(gpu_ctx, rays, monkeypatch)
    with np.errstate(all="ignore"):
        want = O.los_batch(segs, rays, threads=8)
    bad = np.nonzero(fast != want)[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first: {rays[bad[:5]].tolist()}"
    assert np.array_equal(exact, want)
    assert 0 < int(want.sum()) < len(want)


def test_two_pass_los_arbitrary_segments(gpu_ctx, monkeypatch):
    """pm3d_upload_segments accepts any segment list (the Go side hands over g.collisionMap): oblique segments,
    axis-parallel ones off the lattice or with fractional ends, duplicates and points must never be used for a
    first-pass proof they do not support."""
    rng = W.SplitMix64(5150)
    u = rng.uniform(4 * 300).reshape(-1, 4) * 30 + 1
    segs = [u[:60]]                                           # oblique
    a = np.round(u[60:140]);  a[:, 2] = a[:, 0]              # vertical on the lattice, integer ends
    b = np.round(u[140:200]); b[:, 3] = b[:, 1]              # horizontal on the lattice
    c = u[200:240].copy(); c[:, 2] = c[:, 0]                  # vertical off the lattice
    d = np.round(u[240:280]); d[:, 2] = d[:, 0]; d[:, 1] += 0.25; d[:, 3] -= 0.5  # lattice line, fractional ends
    e = np.round(u[280:300]); e[:, 2:] = e[:, :2]             # points
    segs = np.concatenate([segs[0], a, b, c, d, e, a[:10]])
    rays = rng.uniform(4 * 60000).reshape(-1, 4) * 31 + 0.5
    rays[::7] = np.round(rays[::7] * 2) / 2                   # half-lattice endpoints: exact touches
    gpu_ctx.upload_segments(segs)
    fast, exact = _both_modes(gpu_ctx, rays, monkeypatch)
    want = O.los_batch(segs, rays, threads=8)
    assert np.array_equal(exact, want)
    assert np.array_equal(fast, want)
    assert 0 < int(want.sum()) < len(want)
