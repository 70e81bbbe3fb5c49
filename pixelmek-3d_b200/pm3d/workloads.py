"""Synthetic workloads for the five BASELINE.json configs.

Everything here is INPUT generation on the host (numpy): maps, procedural textures, sprite
placements, camera poses and LOS rays, all seeded with splitmix64 so the oracle and the CUDA
path see identical bytes.  Texture PIXELS are procedural (the reference's PNGs are not
copied); the bundled map's GRID, floor grid, lighting and sprite placements come from
tests/golden/*.npz (made by tests/golden/make_fixtures.py from the reference's YAML).
"""
from __future__ import annotations

import math
import os

import numpy as np

from ._lib import SPRITE_DTYPE, VIEW_DTYPE
from .camera import ANCHOR_BOTTOM, CameraState
from .scene import SceneData, premultiply

_M64 = (1 << 64) - 1
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GOLDEN = os.path.join(_ROOT, "tests", "golden")
BOUNDARY_TEX = 32767  # math.MaxInt16, game/model/map.go:458
CLIP_DISTANCE = 0.01  # game/game.go:33
TEX = 256             # resources.TexSize, game/resources/resources.go:35


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _M64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self, n: int) -> np.ndarray:
        """n doubles in [0,1) — vectorised splitmix64 over the counter sequence."""
        idx = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(self.s)
        self.s = int(idx[-1]) if n else self.s
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def _hash2(x: np.ndarray, y: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        v = x.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + y.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(seed)
        v = (v ^ (v >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
        v = v ^ (v >> np.uint64(32))
    return (v & np.uint64(0xFF)).astype(np.int32)


# --------------------------------------------------------------------------- procedural pixels


def wall_tile(seed: int, base=(150, 110, 90), alpha_holes: bool = False) -> np.ndarray:
    """Opaque brick-like tile; with alpha_holes a lattice of fully transparent holes plus a small
    translucent patch (exercises the compose kernel's painter fallback)."""
    yy, xx = np.mgrid[0:TEX, 0:TEX]
    n = _hash2(xx, yy, seed)
    brick = (((yy // 32) % 2) * 32 + xx) % 64
    mortar = ((yy % 32) < 2) | (brick < 2)
    img = np.zeros((TEX, TEX, 4), dtype=np.uint8)
    for ch in range(3):
        c = base[ch] + (n - 128) // 6 + ((xx * (ch + 1) + yy * 3) % 23)
        c = np.where(mortar, 70 + n // 8, c)
        img[..., ch] = np.clip(c, 0, 255)
    img[..., 3] = 255
    if alpha_holes:
        hole = ((xx % 64) > 40) & ((yy % 64) > 40)
        img[hole, 3] = 0
        img[8:16, 8:16, 3] = 128
    return premultiply(img)


def floor_tile(seed: int, base=(60, 120, 50)) -> np.ndarray:
    yy, xx = np.mgrid[0:TEX, 0:TEX]
    n = _hash2(xx // 2, yy // 2, seed)
    img = np.zeros((TEX, TEX, 4), dtype=np.uint8)
    for ch in range(3):
        img[..., ch] = np.clip(base[ch] + (n - 128) // 4 + ((xx ^ yy) & 7), 0, 255)
    img[..., 3] = 255
    return img


def gradient_tile(top, bottom) -> np.ndarray:
    img = np.zeros((TEX, TEX, 4), dtype=np.uint8)
    t = np.linspace(0.0, 1.0, TEX)[:, None]
    for ch in range(3):
        img[..., ch] = np.clip(top[ch] + (bottom[ch] - top[ch]) * t, 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


def sprite_sheet(w: int, h: int, seed: int, columns: int = 1, rows: int = 1) -> np.ndarray:
    """Blob silhouettes (one per sheet cell): opaque body, transparent surround, 1-px translucent rim."""
    yy, xx = np.mgrid[0:h, 0:w]
    cw, ch_ = w / columns, h / rows
    u = (xx % cw) / cw * 2 - 1
    v = (yy % ch_) / ch_ * 2 - 1
    r = np.sqrt(u * u * 1.1 + v * v * 0.9) + 0.08 * np.sin(7 * np.arctan2(v, u) + seed)
    n = _hash2(xx, yy, seed)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    body = r < 0.86
    rim = (r >= 0.86) & (r < 0.90)
    img[..., 0] = np.clip(40 + n // 3 + 60 * (seed % 3), 0, 255)
    img[..., 1] = np.clip(90 + n // 2 - 20 * (seed % 2), 0, 255)
    img[..., 2] = np.clip(40 + n // 4 + 30 * (seed % 5), 0, 255)
    img[..., 3] = np.where(body, 255, np.where(rim, 128, 0))
    return premultiply(img)


# --------------------------------------------------------------------------- maps


def synthetic_levels(w: int, h: int, levels: int, density: float, seed: int, boundary: bool = True, n_tex: int = 4):
    """Random wall cells (texture 1..n_tex) at `density`; a wall's height is uniform in 1..levels."""
    rng = SplitMix64(seed)
    u = rng.uniform(w * h).reshape(w, h)
    t = rng.uniform(w * h).reshape(w, h)
    hgt = rng.uniform(w * h).reshape(w, h)
    grid = np.zeros((levels, w, h), dtype=np.uint16)
    wall = u < density
    tex = (1 + np.floor(t * n_tex)).astype(np.uint16)
    height = (1 + np.floor(hgt * levels)).astype(np.int32)
    for l in range(levels):
        grid[l][wall & (height > l)] = tex[wall & (height > l)]
    if boundary:
        grid[0, 0, :] = grid[0, -1, :] = BOUNDARY_TEX
        grid[0, :, 0] = grid[0, :, -1] = BOUNDARY_TEX
        if levels > 1:
            for l in range(1, levels):
                grid[l, 0, :] = grid[l, -1, :] = 0
                grid[l, :, 0] = grid[l, :, -1] = 0
    return grid


def standard_atlas():
    """tiles: 0..3 walls, 4 side-X variant, 5 side-Y variant, 6 boundary (alpha holes), 7 grass,
    8 path, 9 sky, 10 floor box."""
    tiles = [
        wall_tile(11, (150, 110, 90)), wall_tile(12, (120, 125, 140)), wall_tile(13, (160, 150, 100)),
        wall_tile(14, (110, 100, 120)), wall_tile(15, (90, 140, 150)), wall_tile(16, (170, 120, 120)),
        wall_tile(17, (60, 150, 70), alpha_holes=True), floor_tile(21, (60, 120, 50)), floor_tile(22, (120, 115, 105)),
        gradient_tile((70, 120, 220), (190, 215, 245)), gradient_tile((50, 90, 40), (70, 130, 60)),
    ]
    return np.stack(tiles)


def scene_from_grid(grid: np.ndarray, num_levels: int, floor_grid: np.ndarray | None = None, sheets=None,
                    lighting=(-100.0, 300.0, (0, 0, 0), (255, 255, 255)), has_boundary: bool = True) -> SceneData:
    atlas = standard_atlas()
    # texture table in the style of a map YAML: 1,3,4 plain; 2 has sideX/sideY variants; boundary
    tex_index = np.array([1, 2, 3, 4, BOUNDARY_TEX], dtype=np.int32)
    side0 = np.array([0, 5, 2, 3, 6], dtype=np.int32)  # side 0 -> SideY image when present (map.go:108)
    side1 = np.array([0, 4, 2, 3, 6], dtype=np.int32)  # side 1 -> SideX image when present (map.go:110)
    w, h = grid.shape[1], grid.shape[2]
    ft = np.full((w, h), 7, dtype=np.int16)
    if floor_grid is not None:
        ft[np.asarray(floor_grid) > 0] = 8
    return SceneData(
        grid=grid, num_levels=num_levels, wall_tex_index=tex_index, wall_tile_side0=side0, wall_tile_side1=side1,
        floor_tiles=ft, atlas=atlas, sky_tile=9, floorbox_tile=10, sheets=list(sheets or default_sheets()),
        light_falloff=lighting[0], global_illumination=lighting[1], min_light=tuple(lighting[2]),
        max_light=tuple(lighting[3]), has_boundary_wall=has_boundary,
    )


_SHEETS = None


def default_sheets():
    """Sheet sizes follow the bundled sprites: tree_0 128x198, tree_1 172x224, a 240x42 6x1 mech strip."""
    global _SHEETS
    if _SHEETS is None:
        _SHEETS = [sprite_sheet(128, 198, 1), sprite_sheet(172, 224, 2), sprite_sheet(240, 42, 3, columns=6, rows=1),
                   sprite_sheet(18, 18, 4)]
    return _SHEETS


def load_golden(name: str) -> dict:
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tests/golden/make_fixtures.py where /root/reference exists")
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


_REAL = None


def real_pixels() -> dict:
    """The reference's own wall / floor / sky images and sprite sheets of proving_grounds + trial_day
    (tests/golden/proving_grounds_pixels.npz, decoded from the PNGs by tests/golden/make_pixels.py), premultiplied the
    way the game's loaders hold them."""
    global _REAL
    if _REAL is None:
        z = load_golden("proving_grounds_pixels")
        names = [str(n) for n in z["tile_names"]]
        sheet_names = [str(n) for n in z["sheet_names"]]
        _REAL = dict(tile_names=names, tiles=premultiply(z["tiles"]), sheet_names=sheet_names,
                     sheets=[premultiply(z[f"sheet_{i}"]) for i in range(len(sheet_names))])
    return _REAL


def proving_grounds_scene(pixels: str = "real"):
    """The bundled map.  pixels = "real": the reference's textures and sprite sheets, resolved through the map's own
    texture table the way TextureHandler.TextureAt / MapTexture.GetImage do (game/texture/texture.go:69-97,
    game/model/map.go:106-115: side 0 -> SideY when set, side 1 -> SideX when set, else Image); sheets 0..3 are
    tree_0, tree_1, the adder 6x1 strip and rock_0 — the sizes the procedural sheets imitate.  pixels = "synthetic":
    the procedural tiles and silhouettes (alpha-hole boundary tile, 1-px translucent rims)."""
    g = load_golden("proving_grounds")
    lighting = (float(g["light_falloff"]), float(g["light_illumination"]), tuple(int(v) for v in g["min_light"]),
                tuple(int(v) for v in g["max_light"]))
    if pixels == "synthetic":
        scene = scene_from_grid(g["levels"], int(g["num_raycast_levels"]), g["floor_grid"], lighting=lighting,
                                has_boundary=bool(g["has_boundary_wall"]))
        return scene, g
    rp = real_pixels()
    tile_of = {nm: i for i, nm in enumerate(rp["tile_names"])}
    image = [str(v) for v in g["tex_image"]]
    side_x = [str(v) for v in g["tex_side_x"]]
    side_y = [str(v) for v in g["tex_side_y"]]
    side0 = np.array([tile_of.get(side_y[i] or image[i], -1) for i in range(len(image))], dtype=np.int32)
    side1 = np.array([tile_of.get(side_x[i] or image[i], -1) for i in range(len(image))], dtype=np.int32)
    keep = g["tex_index"] > 0
    floor_images = [str(v) for v in g["floor_images"]]
    ft = np.array([tile_of[nm] for nm in floor_images], dtype=np.int16)[np.asarray(g["floor_grid"], dtype=np.int64)]
    scene = SceneData(
        grid=g["levels"], num_levels=int(g["num_raycast_levels"]), wall_tex_index=g["tex_index"][keep].astype(np.int32),
        wall_tile_side0=side0[keep], wall_tile_side1=side1[keep], floor_tiles=ft, atlas=rp["tiles"],
        sky_tile=tile_of[str(g["sky_box"])], floorbox_tile=tile_of[str(g["floor_box"])], sheets=list(rp["sheets"]),
        light_falloff=lighting[0], global_illumination=lighting[1], min_light=lighting[2], max_light=lighting[3],
        has_boundary_wall=bool(g["has_boundary_wall"]),
    )
    return scene, g


# --------------------------------------------------------------------------- views and sprites


def cardinal_to_angle(compass_degrees: float) -> float:
    """model.CardinalToAngle (game/model/math.go:162-164)."""
    a = math.radians(-compass_degrees) + math.pi / 2
    while a >= 2 * math.pi:
        a -= 2 * math.pi
    while a < 0:
        a += 2 * math.pi
    return a


def make_views(scene: SceneData, w: int, h: int, poses, fov_deg: float = 70.0, render_distance: float = 100.0,
               fov_depth: float = 1.0) -> np.ndarray:
    """poses: iterable of (x, y, z, heading, pitch)."""
    out = []
    st = CameraState(w, h)
    st.fov_angle, st.fov_depth = math.radians(fov_deg), fov_depth
    st.render_distance = render_distance
    st.light_falloff, st.global_illumination = scene.light_falloff, scene.global_illumination
    st.min_light, st.max_light = scene.min_light, scene.max_light
    for x, y, z, heading, pitch in poses:
        st.pos, st.pos_z, st.heading, st.pitch_angle = (float(x), float(y)), float(z), float(heading), float(pitch)
        out.append(st.view_record())
    return np.concatenate(out) if out else np.zeros(0, dtype=VIEW_DTYPE)


def sprite_records(xyzs: np.ndarray, sheet: np.ndarray, anchor: np.ndarray, sheets, illumination=0.0, cell=None):
    """xyzs: [n][4] = x, y, z, scale.  `cell` picks a sheet cell rect via GetSpriteSheetSlices when given."""
    from .camera import sheet_slices

    n = len(xyzs)
    out = np.zeros(n, dtype=SPRITE_DTYPE)
    out["x"], out["y"], out["z"], out["scale"] = xyzs[:, 0], xyzs[:, 1], xyzs[:, 2], xyzs[:, 3]
    out["illumination"] = illumination
    out["sheet"] = sheet
    out["anchor"] = anchor
    out["focusable"] = 1
    for i in range(n):
        s = sheets[int(sheet[i])]
        sh, sw = s.shape[0], s.shape[1]
        if int(sheet[i]) == 2:  # the 6x1 strip: use a cell rect (1 px short, as the reference slices them)
            rects = sheet_slices(sw, sh, 6, 1)
            r = rects[(i if cell is None else int(cell[i])) % 6]
        else:
            r = (0, 0, sw, sh)
        out["rx0"][i], out["ry0"][i], out["rx1"][i], out["ry1"][i] = r
    return out


def frustum_sprites(scene: SceneData, pose, n: int, seed: int, max_dist: float = 100.0, fov_deg: float = 70.0):
    """n sprites uniform in the view frustum within max_dist cells, scale U[0.3,1], anchor bottom (C2)."""
    rng = SplitMix64(seed)
    x, y, _, heading, _ = pose
    mw, mh = scene.map_size
    # uniform over (frustum within max_dist) ∩ map: candidates outside the map are rejected, not clamped onto its edge
    # (clamping piles hundreds of billboards onto one screen line whenever the camera faces a nearby map border)
    m = 16 * n
    sx, sy = np.zeros(0), np.zeros(0)
    for _ in range(32):
        r = 1.0 + (max_dist - 1.0) * np.sqrt(rng.uniform(m))
        a = heading + (rng.uniform(m) - 0.5) * math.radians(fov_deg)
        cx, cy = x + r * np.cos(a), y + r * np.sin(a)
        ok = (cx >= 1.01) & (cx <= mw - 1.01) & (cy >= 1.01) & (cy <= mh - 1.01)
        sx, sy = np.concatenate([sx, cx[ok]]), np.concatenate([sy, cy[ok]])
        if len(sx) >= n:
            break
    if len(sx) < n:  # the frustum barely touches the map: the rest goes anywhere within max_dist (mostly off screen)
        k = n - len(sx)
        sx = np.concatenate([sx, np.clip(x + (rng.uniform(k) * 2 - 1) * max_dist, 1.01, mw - 1.01)])
        sy = np.concatenate([sy, np.clip(y + (rng.uniform(k) * 2 - 1) * max_dist, 1.01, mh - 1.01)])
    sx, sy = sx[:n], sy[:n]
    scale = 0.3 + 0.7 * rng.uniform(n)
    sheet = (np.floor(rng.uniform(n) * 3)).astype(np.int32)
    xyzs = np.stack([sx, sy, np.zeros(n), scale], axis=1)
    return sprite_records(xyzs, sheet, np.full(n, ANCHOR_BOTTOM, np.uint8), scene.sheets)


def golden_sprites(scene: SceneData, g: dict) -> np.ndarray:
    """The bundled map's static sprites (tree_0 -> sheet 0, tree_1 -> sheet 1) + mission units as mech cells."""
    xyzs = g["sprite_xyzs"]
    names = [str(s) for s in g["sprite_images"]]
    sheet_of = np.array([0 if "tree_0" in nm else 1 for nm in names], dtype=np.int32)
    sheet = sheet_of[g["sprite_image"]]
    recs = sprite_records(xyzs, sheet, g["sprite_anchor"], scene.sheets)
    if "mission_units" in g and len(g["mission_units"]):
        mu = g["mission_units"]
        u = np.stack([mu[:, 0], mu[:, 1], mu[:, 2], np.full(len(mu), 0.75)], axis=1)
        recs = np.concatenate([recs, sprite_records(u, np.full(len(mu), 2, np.int32), np.zeros(len(mu), np.uint8), scene.sheets)])
    return recs


def free_cell_poses(scene: SceneData, n: int, seed: int, z: float = 0.5):
    """n poses uniform over free level-0 cells (cell centre jittered), heading U[0,2pi), pitch 0."""
    rng = SplitMix64(seed)
    free = np.argwhere(scene.grid[0] == 0)
    pick = (rng.uniform(n) * len(free)).astype(np.int64)
    jx, jy = rng.uniform(n), rng.uniform(n)
    hd = rng.uniform(n) * 2 * math.pi
    cells = free[pick]
    return [(cells[i, 0] + 0.05 + 0.9 * jx[i], cells[i, 1] + 0.05 + 0.9 * jy[i], z, hd[i], 0.0) for i in range(n)]


def nearest_sprites(all_sprites: np.ndarray, poses, k: int):
    """Per view the k nearest sprites (C4: '50 nearest sprites each'); returns (records, offsets)."""
    xs, ys = all_sprites["x"], all_sprites["y"]
    recs, off = [], [0]
    for p in poses:
        d = (xs - p[0]) ** 2 + (ys - p[1]) ** 2
        idx = np.argsort(d, kind="stable")[:k]
        recs.append(all_sprites[np.sort(idx)])
        off.append(off[-1] + len(idx))
    return (np.concatenate(recs) if recs else np.zeros(0, SPRITE_DTYPE)), np.array(off, dtype=np.int32)


# --------------------------------------------------------------------------- the five configs


def config_c1(pixels: str = "real"):
    """Bundled proving_grounds, trial_day drop zone, 640x360, static sprites + mission units."""
    scene, g = proving_grounds_scene(pixels)
    pose = (float(g["drop_zone"][0]), float(g["drop_zone"][1]), 0.5, cardinal_to_angle(float(g["drop_heading_deg"])), 0.0)
    views = make_views(scene, 640, 360, [pose])
    return dict(scene=scene, w=640, h=360, views=views, sprites=golden_sprites(scene, g), poses=[pose], golden=g)


def config_c2(w: int = 1920, h: int = 1080, n_sprites: int = 200, n_poses: int = 1, seed: int = 12345, pixels: str = "real"):
    """Same map at 1080p with 200 frustum sprites; n_poses > 1 gives a ring of distinct views."""
    scene, g = proving_grounds_scene(pixels)
    base = (float(g["drop_zone"][0]), float(g["drop_zone"][1]), 0.5, cardinal_to_angle(float(g["drop_heading_deg"])), 0.0)
    poses = [base]
    if n_poses > 1:
        extra = free_cell_poses(scene, n_poses - 1, seed + 1)
        # keep them in the built-up corner of the map so walls, paths and sprites stay in view
        rng = SplitMix64(seed + 2)
        u = rng.uniform(2 * (n_poses - 1))
        poses += [(2.0 + 56.0 * u[2 * i], 2.0 + 56.0 * u[2 * i + 1], 0.5, e[3], 0.0) for i, e in enumerate(extra)]
        poses = [p if scene.grid[0][int(p[0])][int(p[1])] == 0 else base for p in poses]
    views = make_views(scene, w, h, poses)
    recs, off = [], [0]
    for i, p in enumerate(poses):
        r = frustum_sprites(scene, p, n_sprites, seed + 17 * i)
        recs.append(r)
        off.append(off[-1] + len(r))
    return dict(scene=scene, w=w, h=h, views=views, sprites=np.concatenate(recs), offsets=np.array(off, np.int32), poses=poses)


def config_c3(levels: int = 3, w: int = 3840, h: int = 2160, pitch: float = 0.0, seed: int = 12345, map_size: int = 512):
    grid = synthetic_levels(map_size, map_size, min(levels, 3), 0.05, seed)
    scene = scene_from_grid(grid, levels)
    free = np.argwhere(grid[0] == 0)
    c = free[len(free) // 2 + 7]
    pose = (c[0] + 0.5, c[1] + 0.5, 0.5, 0.6, pitch)
    views = make_views(scene, w, h, [pose], render_distance=-1.0)
    return dict(scene=scene, w=w, h=h, views=views, sprites=np.zeros(0, SPRITE_DTYPE), poses=[pose])


def config_c4(n_views: int = 4096, w: int = 320, h: int = 200, k_sprites: int = 50, seed: int = 4096, pixels: str = "real"):
    scene, g = proving_grounds_scene(pixels)
    poses = free_cell_poses(scene, n_views, seed)
    views = make_views(scene, w, h, poses)
    sprites, off = nearest_sprites(golden_sprites(scene, g), poses, k_sprites)
    return dict(scene=scene, w=w, h=h, views=views, sprites=sprites, offsets=off, poses=poses)


def config_c5(n_rays: int = 1 << 24, map_size: int = 256, density: float = 0.10, seed: int = 42):
    """256x256x1 map, 10 % walls + boundary ring -> collision segments; rays between free cells."""
    from .context import generate_wall_collision_lines

    grid = synthetic_levels(map_size, map_size, 1, density, seed)
    segs = generate_wall_collision_lines(grid[0], True, CLIP_DISTANCE)
    rays = los_rays(grid[0], n_rays, seed + 1)
    return dict(grid=grid, segments=segs, rays=rays)


def los_rays(level0: np.ndarray, n: int, seed: int) -> np.ndarray:
    rng = SplitMix64(seed)
    free = np.argwhere(level0 == 0)
    a = (rng.uniform(n) * len(free)).astype(np.int64)
    b = (rng.uniform(n) * len(free)).astype(np.int64)
    j = rng.uniform(4 * n).reshape(n, 4)
    rays = np.empty((n, 4), dtype=np.float64)
    rays[:, 0] = free[a, 0] + 0.05 + 0.9 * j[:, 0]
    rays[:, 1] = free[a, 1] + 0.05 + 0.9 * j[:, 1]
    rays[:, 2] = free[b, 0] + 0.05 + 0.9 * j[:, 2]
    rays[:, 3] = free[b, 1] + 0.05 + 0.9 * j[:, 3]
    return rays


def path_queries(level0: np.ndarray, n: int, seed: int, max_manhattan: int | None = None) -> np.ndarray:
    """(start x, y, finish x, y) between free cells, fractional positions inside the cells; optionally no farther
    apart than max_manhattan cells (AI units path towards targets in their neighbourhood)."""
    rng = SplitMix64(seed)
    free = np.argwhere(level0 == 0)
    a = free[(rng.uniform(n) * len(free)).astype(np.int64)]
    b = free[(rng.uniform(n) * len(free)).astype(np.int64)]
    if max_manhattan is not None:
        for _ in range(64):
            far = np.abs(a - b).sum(axis=1) > max_manhattan
            if not far.any():
                break
            b[far] = free[(rng.uniform(int(far.sum())) * len(free)).astype(np.int64)]
        far = np.abs(a - b).sum(axis=1) > max_manhattan
        b[far] = a[far]
    j = rng.uniform(4 * n).reshape(n, 4)
    q = np.empty((n, 4), dtype=np.float64)
    q[:, :2] = a + 0.05 + 0.9 * j[:, :2]
    q[:, 2:] = b + 0.05 + 0.9 * j[:, 2:]
    return q


def move_workload(level0: np.ndarray, n_movers: int, n_colliders: int, seed: int, max_len: float = 0.5):
    """A tick of the projectile / unit fan-out (game/combat.go:206-239): colliders scattered over free cells,
    movers (a third of them next to a collider) each asking for one short move.  Returns (colliders, movers)."""
    from ._lib import COLLIDER_DTYPE, MOVER_DTYPE

    rng = SplitMix64(seed)
    free = np.argwhere(level0 == 0)
    c = np.zeros(n_colliders, dtype=COLLIDER_DTYPE)
    cells = free[(rng.uniform(n_colliders) * len(free)).astype(np.int64)]
    c["x"] = cells[:, 0] + rng.uniform(n_colliders)
    c["y"] = cells[:, 1] + rng.uniform(n_colliders)
    c["radius"] = 0.1 + 0.4 * rng.uniform(n_colliders)
    c["height"] = 0.3 + rng.uniform(n_colliders)
    m = np.zeros(n_movers, dtype=MOVER_DTYPE)
    cells = free[(rng.uniform(n_movers) * len(free)).astype(np.int64)]
    m["pos_x"] = cells[:, 0] + rng.uniform(n_movers)
    m["pos_y"] = cells[:, 1] + rng.uniform(n_movers)
    if n_colliders:
        near = rng.uniform(n_movers) < 0.33
        pick = (rng.uniform(n_movers) * n_colliders).astype(np.int64)
        m["pos_x"] = np.where(near, c["x"][pick] + (rng.uniform(n_movers) - 0.5) * 2.0, m["pos_x"])
        m["pos_y"] = np.where(near, c["y"][pick] + (rng.uniform(n_movers) - 0.5) * 2.0, m["pos_y"])
    m["pos_z"] = 0.3 * rng.uniform(n_movers)
    ang = rng.uniform(n_movers) * (2.0 * math.pi)
    length = rng.uniform(n_movers) * max_len
    m["move_x"] = m["pos_x"] + np.cos(ang) * length
    m["move_y"] = m["pos_y"] + np.sin(ang) * length
    m["move_z"] = m["pos_z"]
    m["radius"] = 0.01
    m["height"] = 0.01
    m["self"] = -1
    m["parent"] = -1
    m["anchor"] = 1
    return c, m
