"""ctypes front end of the CPU ORACLE (oracle/pm3d_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/pm3d_oracle.h.  PARITY UNPINNED against the Go binary.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under pixelmek-3d_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SPRITE_DTYPE = np.dtype(
    [
        ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("scale", "<f8"), ("illumination", "<f8"),
        ("sheet", "<i4"), ("rx0", "<i4"), ("ry0", "<i4"), ("rx1", "<i4"), ("ry1", "<i4"),
        ("anchor", "u1"), ("focusable", "u1"), ("_pad", "u1", (2,)),
    ],
    align=False,
)
assert SPRITE_DTYPE.itemsize == 64

COLUMN_DTYPE = np.dtype(
    [
        ("perp_dist", "<f8"), ("line_height", "<i8"), ("draw_start", "<i8"), ("draw_end", "<i8"),
        ("map_x", "<i4"), ("map_y", "<i4"), ("tex_x", "<i4"), ("tile", "<i2"), ("side", "u1"), ("hit", "u1"),
        ("tint", "u1", (4,)), ("_tail", "u1", (4,)),
    ],
    align=False,
)
assert COLUMN_DTYPE.itemsize == 56


class OrcScene(C.Structure):
    _fields_ = [
        ("map_w", C.c_int32), ("map_h", C.c_int32), ("stored_levels", C.c_int32), ("num_levels", C.c_int32),
        ("grid", C.c_void_p), ("wall_lut", C.c_void_p), ("floor_tiles", C.c_void_p), ("atlas", C.c_void_p),
        ("n_tiles", C.c_int32), ("tex_size", C.c_int32), ("sky_tile", C.c_int32), ("floorbox_tile", C.c_int32),
        ("sheet_pixels", C.c_void_p), ("sheet_offset", C.c_void_p), ("sheet_w", C.c_void_p), ("sheet_h", C.c_void_p),
        ("n_sheets", C.c_int32), ("_pad", C.c_int32),
    ]


class OrcCamera(C.Structure):
    _fields_ = [
        ("pos_x", C.c_double), ("pos_y", C.c_double), ("cam_z_px", C.c_double),
        ("dir_x", C.c_double), ("dir_y", C.c_double), ("plane_x", C.c_double), ("plane_y", C.c_double),
        ("pitch_px", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("_pad", C.c_int32),
        ("render_distance", C.c_double), ("light_falloff", C.c_double), ("global_illumination", C.c_double),
        ("min_light", C.c_uint8 * 4), ("max_light", C.c_uint8 * 4),
    ]


class OrcConvergence(C.Structure):
    _fields_ = [("distance", C.c_double), ("point_x", C.c_double), ("point_y", C.c_double),
                ("sprite", C.c_int32), ("_pad", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile oracle/libpm3d_oracle.so with the committed Makefile (building the checker is
    not using it)."""
    so = os.path.join(_HERE, "libpm3d_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpm3d_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_go_int.restype = C.c_int64
        _LIB.orc_go_int.argtypes = [C.c_double]
        _LIB.orc_line_intersection.restype = C.c_int
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """Flattened Map + TextureHandler + sprite sheets.  `data` is any object with the
    attributes of pm3d.scene.SceneData (duck-typed so oracle/ does not import the product)."""

    def __init__(self, data):
        self.grid = np.ascontiguousarray(data.grid, dtype=np.uint16)
        ls, w, h = self.grid.shape
        self.wall_lut = np.full((32768, 2), -1, dtype=np.int16)
        for tex, t0, t1 in zip(data.wall_tex_index, data.wall_tile_side0, data.wall_tile_side1):
            self.wall_lut[int(tex), 0] = t0
            self.wall_lut[int(tex), 1] = t1
        self.floor_tiles = np.ascontiguousarray(data.floor_tiles, dtype=np.int16)
        assert self.floor_tiles.shape == (w, h)
        self.atlas = np.ascontiguousarray(data.atlas, dtype=np.uint8)
        nt, ts, ts2, ch = self.atlas.shape
        assert ts == ts2 and ch == 4
        sheets = [np.ascontiguousarray(s, dtype=np.uint8) for s in data.sheets]
        sizes = [s.size for s in sheets]
        self.sheet_offset = np.zeros(max(len(sheets), 1), dtype=np.int64)
        if sheets:
            self.sheet_offset[: len(sheets)] = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        self.sheet_pixels = np.concatenate([s.reshape(-1) for s in sheets]) if sheets else np.zeros(4, np.uint8)
        self.sheet_w = np.array([s.shape[1] for s in sheets] or [0], dtype=np.int32)
        self.sheet_h = np.array([s.shape[0] for s in sheets] or [0], dtype=np.int32)
        self.c = OrcScene(
            w, h, ls, int(data.num_levels), _ptr(self.grid), _ptr(self.wall_lut), _ptr(self.floor_tiles),
            _ptr(self.atlas), nt, ts, int(data.sky_tile), int(data.floorbox_tile), _ptr(self.sheet_pixels),
            _ptr(self.sheet_offset), _ptr(self.sheet_w), _ptr(self.sheet_h), len(sheets), 0,
        )


def make_camera(p) -> OrcCamera:
    """p: any object/dict-like with the low-level pose fields (pm3d.camera.Camera.pose())."""
    g = (lambda k: p[k]) if isinstance(p, dict) else (lambda k: getattr(p, k))
    cam = OrcCamera()
    for k in ("pos_x", "pos_y", "cam_z_px", "dir_x", "dir_y", "plane_x", "plane_y", "render_distance",
              "light_falloff", "global_illumination"):
        setattr(cam, k, float(g(k)))
    for k in ("pitch_px", "w", "h"):
        setattr(cam, k, int(g(k)))
    mn, mx = g("min_light"), g("max_light")
    for i in range(3):
        cam.min_light[i] = int(mn[i])
        cam.max_light[i] = int(mx[i])
    cam.min_light[3] = cam.max_light[3] = 255
    return cam


def cast_walls(scene: Scene, cam: OrcCamera):
    cols = np.zeros((scene.c.num_levels, cam.w), dtype=COLUMN_DTYPE)
    zbuf = np.zeros(cam.w, dtype=np.float64)
    lib().orc_cast_walls(C.byref(scene.c), C.byref(cam), _ptr(cols), _ptr(zbuf))
    return cols, zbuf


def render(scene: Scene, cam: OrcCamera, sprites: np.ndarray, threads: int = 1):
    sprites = np.ascontiguousarray(sprites, dtype=SPRITE_DTYPE)
    n = len(sprites)
    rgba = np.zeros((cam.h, cam.w, 4), dtype=np.uint8)
    rects = np.zeros((max(n, 1), 4), dtype=np.int32)
    conv = OrcConvergence()
    cols = np.zeros((scene.c.num_levels, cam.w), dtype=COLUMN_DTYPE)
    zbuf = np.zeros(cam.w, dtype=np.float64)
    lib().orc_render(C.byref(scene.c), C.byref(cam), _ptr(sprites), C.c_int32(n), _ptr(rgba), _ptr(rects),
                     C.byref(conv), _ptr(cols), _ptr(zbuf), C.c_int32(threads))
    return {
        "rgba": rgba, "rects": rects[:n], "columns": cols, "zbuf": zbuf,
        "convergence": {"distance": conv.distance, "point": (conv.point_x, conv.point_y), "sprite": conv.sprite},
    }


def sprite_order(cam: OrcCamera, sprites: np.ndarray):
    sprites = np.ascontiguousarray(sprites, dtype=SPRITE_DTYPE)
    order = np.zeros(max(len(sprites), 1), dtype=np.int32)
    lib().orc_sprite_order(C.byref(cam), _ptr(sprites), C.c_int32(len(sprites)), _ptr(order))
    return order[: len(sprites)]


def line_intersection(l1, l2):
    a = np.ascontiguousarray(l1, dtype=np.float64)
    b = np.ascontiguousarray(l2, dtype=np.float64)
    px, py = C.c_double(), C.c_double()
    ok = lib().orc_line_intersection(_ptr(a), _ptr(b), C.byref(px), C.byref(py))
    return (px.value, py.value, True) if ok else (0.0, 0.0, False)


def los_batch(segs: np.ndarray, rays: np.ndarray, threads: int = 1) -> np.ndarray:
    segs = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 4)
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 4)
    out = np.zeros(len(rays), dtype=np.uint8)
    lib().orc_los_batch(_ptr(segs), C.c_int32(len(segs)), _ptr(rays), C.c_int64(len(rays)), _ptr(out),
                        C.c_int32(threads))
    return out


def move_walls_batch(scene: Scene, segs: np.ndarray, moves: np.ndarray, clip: float, threads: int = 1):
    segs = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 4)
    moves = np.ascontiguousarray(moves, dtype=np.float64).reshape(-1, 4)
    n = len(moves)
    pos = np.zeros((n, 2), dtype=np.float64)
    collided = np.zeros(n, dtype=np.uint8)
    hits = np.zeros(n, dtype=np.int32)
    lib().orc_move_walls_batch(C.byref(scene.c), _ptr(segs), C.c_int32(len(segs)), _ptr(moves), C.c_int64(n),
                               C.c_double(clip), _ptr(pos), _ptr(collided), _ptr(hits), C.c_int32(threads))
    return pos, collided, hits


MOVER_DTYPE = np.dtype(
    [
        ("pos_x", "<f8"), ("pos_y", "<f8"), ("pos_z", "<f8"), ("move_x", "<f8"), ("move_y", "<f8"), ("move_z", "<f8"),
        ("radius", "<f8"), ("height", "<f8"), ("self", "<i4"), ("parent", "<i4"),
        ("anchor", "u1"), ("destroyed", "u1"), ("_pad", "u1", (6,)),
    ]
)
assert MOVER_DTYPE.itemsize == 80
COLLIDER_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("radius", "<f8"), ("height", "<f8"),
     ("anchor", "u1"), ("destroyed", "u1"), ("_pad", "u1", (6,))]
)
assert COLLIDER_DTYPE.itemsize == 48
MOVE_RESULT_DTYPE = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("collided", "<i4"), ("n_collisions", "<i4"), ("first_entity", "<i4"),
     ("n_points", "<i4"), ("first_x", "<f8"), ("first_y", "<f8"), ("first_z", "<f8"), ("near_x", "<f8"), ("near_y", "<f8")]
)
assert MOVE_RESULT_DTYPE.itemsize == 80


def line_circle_intersection(line, cx, cy, radius, is_segment=True):
    a = np.ascontiguousarray(line, dtype=np.float64)
    out = np.zeros(4, dtype=np.float64)
    n = lib().orc_line_circle_intersection(_ptr(a), C.c_double(cx), C.c_double(cy), C.c_double(radius),
                                           C.c_int(int(is_segment)), _ptr(out))
    return [(float(out[2 * i]), float(out[2 * i + 1])) for i in range(n)]


def valid_move_batch(scene: Scene, segs, colliders, player: int, clip: float, movers, threads: int = 1):
    """Game.getValidMove(..., checkAlternate=false) for a batch of movers (game/collision.go:69-203, 307-355)."""
    segs = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 4)
    colliders = np.ascontiguousarray(colliders, dtype=COLLIDER_DTYPE).reshape(-1)
    movers = np.ascontiguousarray(movers, dtype=MOVER_DTYPE).reshape(-1)
    out = np.zeros(len(movers), dtype=MOVE_RESULT_DTYPE)
    lib().orc_valid_move_batch(C.byref(scene.c), _ptr(segs), C.c_int32(len(segs)), _ptr(colliders),
                               C.c_int32(len(colliders)), C.c_int32(player), C.c_double(clip), _ptr(movers),
                               C.c_int64(len(movers)), _ptr(out), C.c_int32(threads))
    return out


def valid_move_alternate(scene: Scene, segs, colliders, player: int, clip: float, mover, heading: float,
                         velocity: float, check_alternate: bool = True):
    """Game.getValidMove with the checkAlternate escape heuristics for ONE mover (game/collision.go:204-306)."""
    segs = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 4)
    colliders = np.ascontiguousarray(colliders, dtype=COLLIDER_DTYPE).reshape(-1)
    mover = np.ascontiguousarray(mover, dtype=MOVER_DTYPE).reshape(-1)[:1]
    out = np.zeros(1, dtype=MOVE_RESULT_DTYPE)
    lib().orc_valid_move_alternate(C.byref(scene.c), _ptr(segs), C.c_int32(len(segs)), _ptr(colliders),
                                   C.c_int32(len(colliders)), C.c_int32(player), C.c_double(clip), _ptr(mover),
                                   C.c_double(heading), C.c_double(velocity), C.c_int(int(check_alternate)), _ptr(out))
    return out[0]


def present(src: np.ndarray, out_w: int, out_h: int, render_scale: float, mode: int, threads: int = 4) -> np.ndarray:
    """rayScreen -> overlayScreen (game/scene_game.go:169-185): 1/renderScale nearest scale + copy / CRT shader."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    sh, sw = src.shape[:2]
    out = np.zeros((out_h, out_w, 4), dtype=np.uint8)
    lib().orc_present(_ptr(src), C.c_int32(sw), C.c_int32(sh), C.c_int32(out_w), C.c_int32(out_h),
                      C.c_double(render_scale), C.c_int32(mode), _ptr(out), C.c_int32(threads))
    return out


def go_int(v: float) -> int:
    return int(lib().orc_go_int(float(v)))


# ---------------------------------------------------------------- Go math/rand + ClutterHandler.Update

CLUTTER_DEF_DTYPE = np.dtype([("frequency", "<f8"), ("scale", "<f8"), ("sheet", "<i4"), ("rx0", "<i4"), ("ry0", "<i4"),
                              ("rx1", "<i4"), ("ry1", "<i4"), ("_pad", "<i4")])
CLUTTER_QUERY_DTYPE = np.dtype([("cam_x", "<f8"), ("cam_y", "<f8"), ("n_angles", "<i4"), ("_pad", "<i4"),
                                ("cs", "<f8", (21, 2))])
CLUTTER_SPRITE_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("cell_x", "<i4"), ("cell_y", "<i4"), ("def", "<i4"),
                                 ("_pad", "<i4")])
assert CLUTTER_DEF_DTYPE.itemsize == 40 and CLUTTER_QUERY_DTYPE.itemsize == 360 and CLUTTER_SPRITE_DTYPE.itemsize == 32


class _GoRandState(C.Structure):
    _fields_ = [("tap", C.c_int32), ("feed", C.c_int32), ("vec", C.c_uint64 * 607)]


class GoRand:
    """rand.New(rand.NewSource(seed)) of Go's math/rand (oracle/pm3d_oracle_clutter.c)."""

    def __init__(self, seed: int):
        self._s = _GoRandState()
        self.seed(seed)

    def seed(self, seed: int):
        lib().orc_gorand_seed(C.byref(self._s), C.c_int64(seed))

    def uint64(self) -> int:
        lib().orc_gorand_uint64.restype = C.c_uint64
        return int(lib().orc_gorand_uint64(C.byref(self._s)))

    def int63(self) -> int:
        lib().orc_gorand_int63.restype = C.c_int64
        return int(lib().orc_gorand_int63(C.byref(self._s)))

    def intn(self, n: int) -> int:
        lib().orc_gorand_intn.restype = C.c_int32
        return int(lib().orc_gorand_intn(C.byref(self._s), C.c_int32(n)))

    def float64(self) -> float:
        lib().orc_gorand_float64.restype = C.c_double
        return float(lib().orc_gorand_float64(C.byref(self._s)))


def gorand_cooked() -> np.ndarray:
    """The rngCooked table of math/rand (607 words), regenerated the way gen_cooked.go makes it."""
    lib().orc_gorand_cooked.restype = C.POINTER(C.c_uint64)
    return np.ctypeslib.as_array(lib().orc_gorand_cooked(), shape=(607,)).copy()


def gorand_gen_state(n: int, jump: bool) -> np.ndarray:
    out = np.zeros(607, dtype=np.uint64)
    lib().orc_gorand_gen_state(C.c_uint64(n), C.c_int32(int(jump)), _ptr(out))
    return out


def clutter_angles(heading: float, fov_degrees: float) -> np.ndarray:
    """The sample angles of game/clutter.go:44-45."""
    out = np.zeros(64, dtype=np.float64)
    n = lib().orc_clutter_angles(C.c_double(heading), C.c_double(fov_degrees), _ptr(out), C.c_int32(64))
    if n < 0:
        raise ValueError("fov must be positive")
    return out[:n]


def clutter_update(level0: np.ndarray, cell_class: np.ndarray, class_mask: np.ndarray, defs: np.ndarray, map_seed: int,
                   clutter_distance: float, query: np.ndarray) -> np.ndarray:
    """ClutterHandler.Update(g, true) on an empty handler for ONE camera (game/clutter.go:25-117)."""
    level0 = np.ascontiguousarray(level0, dtype=np.uint16)
    map_w, map_h = level0.shape
    cell_class = np.ascontiguousarray(cell_class, dtype=np.uint8).reshape(map_w, map_h)
    class_mask = np.ascontiguousarray(class_mask, dtype=np.uint64)
    defs = np.ascontiguousarray(defs, dtype=CLUTTER_DEF_DTYPE).reshape(-1)
    query = np.ascontiguousarray(query, dtype=CLUTTER_QUERY_DTYPE).reshape(-1)[:1]
    cap = 21 * (int(max(clutter_distance, 0)) + 2) * max(len(defs), 1)
    out = np.zeros(cap, dtype=CLUTTER_SPRITE_DTYPE)
    f = lib().orc_clutter_update
    f.restype = C.c_int64
    n = f(_ptr(level0), C.c_int32(map_w), C.c_int32(map_h), _ptr(cell_class), _ptr(class_mask), _ptr(defs),
          C.c_int32(len(defs)), C.c_int64(map_seed), C.c_double(clutter_distance), _ptr(query), _ptr(out), C.c_int64(cap))
    assert 0 <= n <= cap
    return out[:n]


def raycast_sprites(world: np.ndarray, cam_x: float, cam_y: float, render_distance: float, target: int = -1,
                    clutter: np.ndarray | None = None, extras: np.ndarray | None = None):
    """Game.getRaycastSprites for ONE camera (game/sprites.go:102-145): (sprites, index)."""
    world = np.ascontiguousarray(world, dtype=SPRITE_DTYPE).reshape(-1)
    clutter = np.zeros(0, SPRITE_DTYPE) if clutter is None else np.ascontiguousarray(clutter, dtype=SPRITE_DTYPE).reshape(-1)
    extras = np.zeros(0, SPRITE_DTYPE) if extras is None else np.ascontiguousarray(extras, dtype=SPRITE_DTYPE).reshape(-1)
    cap = len(world) + len(clutter) + len(extras)
    out = np.zeros(max(cap, 1), dtype=SPRITE_DTYPE)
    idx = np.zeros(max(cap, 1), dtype=np.int32)
    f = lib().orc_raycast_sprites
    f.restype = C.c_int32
    n = f(_ptr(world), C.c_int32(len(world)), C.c_double(cam_x), C.c_double(cam_y), C.c_double(render_distance),
          C.c_int32(target), _ptr(clutter), C.c_int32(len(clutter)), _ptr(extras), C.c_int32(len(extras)), _ptr(out), _ptr(idx),
          C.c_int32(cap))
    return out[:n], idx[:n]


def find_path_batch(level0: np.ndarray, queries: np.ndarray, cap: int = 4096, threads: int = 8):
    """Pathing.FindPath for a batch of (start x, y, finish x, y) queries: (points [n][cap][2], count [n] (-1 = no
    path), astar path cells [n], nodes popped [n])."""
    L = lib()
    level0 = np.ascontiguousarray(level0, dtype=np.uint16)
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 4)
    n = len(q)
    pts = np.zeros((n, cap, 2), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    cells = np.zeros(n, dtype=np.int32)
    exp = np.zeros(n, dtype=np.int64)
    L.orc_find_path_batch.restype = None
    L.orc_find_path_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.orc_find_path_batch(_ptr(level0), level0.shape[0], level0.shape[1], _ptr(q), n, _ptr(pts), cap, _ptr(cnt), _ptr(cells),
                          _ptr(exp), threads)
    return pts, cnt, cells, exp
