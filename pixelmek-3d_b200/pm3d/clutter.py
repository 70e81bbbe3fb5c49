"""Host-side mirror of the reference's clutter boundary, in Python because no Go toolchain
exists in this image.  Same names and argument meaning as the Go side:

  model.MapClutter                           game/model/map.go:128-133
  NewClutterHandler / (*ClutterHandler).Update(g, forceUpdate)   game/clutter.go:16-117

Which cells are in view, which of them are new, the Go math/rand draws that place a cell's rocks and shrubs — all
of that runs on the device (pm3d_clutter_batch).  What stays here is what the Go shim does once per mission:
running the FloorPathMatch regular expressions over the distinct floor texture paths, and the angle loop of
clutter.go:44-45 with its cos/sin (host transcendentals, as everywhere in this package).  Nothing here draws a
random number or walks a ray on the CPU.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass

import numpy as np

from ._lib import CLUTTER_DEF_DTYPE, CLUTTER_MAX_ANGLES, CLUTTER_QUERY_DTYPE, SPRITE_DTYPE
from .context import Context

METERS_PER_UNIT = 20.0  # game/model/model.go:12
ANCHOR_BOTTOM = 0


@dataclass
class MapClutter:
    """model.MapClutter (game/model/map.go:128-133)"""
    image: str
    frequency: float
    height: float
    floorPathMatch: str | None = None


def sample_angles(cam_heading: float, fov_degrees: float) -> list[float]:
    """clutter.go:44-45: for a := camHeading - viewFOV/2; a <= camHeading+viewFOV/2; a += viewFOV / 20.
    (21 angles, or 20 when the accumulated sum ends a rounding error above the limit.)"""
    view_fov = fov_degrees * math.pi / 180  # geom.Radians
    step = view_fov / 20
    if not step > 0:
        raise ValueError("fovDegrees must be positive")
    out = []
    a = cam_heading - view_fov / 2
    while a <= cam_heading + view_fov / 2:
        out.append(a)
        a += step
    assert len(out) <= CLUTTER_MAX_ANGLES
    return out


def clutter_queries(cameras, fov_degrees: float) -> np.ndarray:
    """pm3d_clutter_query records for cameras = iterable of (camX, camY, camHeading)."""
    cams = list(cameras)
    q = np.zeros(len(cams), dtype=CLUTTER_QUERY_DTYPE)
    for i, (x, y, heading) in enumerate(cams):
        ang = sample_angles(float(heading), fov_degrees)
        q["cam_x"][i], q["cam_y"][i], q["n_angles"][i] = x, y, len(ang)
        for k, a in enumerate(ang):
            q["cs"][i, k, 0], q["cs"][i, k, 1] = math.cos(a), math.sin(a)
    return q


def resolve_clutter(clutter: list[MapClutter], floor_class: np.ndarray, class_paths: list[str], sheet_of_image: dict):
    """What the shim resolves once per mission: pm3d_clutter_def records and, per distinct floor texture path, the mask
    of clutter entries whose FloorPathMatch accepts it (clutter.go:79-81: nil matches everything).
    sheet_of_image: clutter image path -> (sheet id, width, height)."""
    defs = np.zeros(len(clutter), dtype=CLUTTER_DEF_DTYPE)
    for i, c in enumerate(clutter):
        sheet, w, h = sheet_of_image[c.image]
        defs["frequency"][i] = c.frequency
        defs["scale"][i] = c.height / METERS_PER_UNIT                      # clutter.go:88
        defs["sheet"][i] = sheet
        defs["rx0"][i], defs["ry0"][i], defs["rx1"][i], defs["ry1"][i] = 0, 0, w - 1, h - 1  # sprite.go:176
    masks = np.zeros(len(class_paths), dtype=np.uint64)
    for c, path in enumerate(class_paths):
        m = 0
        for i, cl in enumerate(clutter):
            if cl.floorPathMatch is None or re.search(cl.floorPathMatch, path):
                m |= 1 << i
        masks[c] = m
    return defs, np.ascontiguousarray(floor_class, dtype=np.uint8), masks


class ClutterHandler:
    """game/clutter.go:10-23.  `sprites` / `spritesByPosition` hold what the Go maps hold after Update: the clutter
    sprites of the cells currently in view (pm3d_sprite records, ready for the camera's sprite list, as
    getRaycastSprites appends them: game/sprites.go:121-124)."""

    def __init__(self, ctx: Context, clutter: list[MapClutter], floor_class: np.ndarray, class_paths: list[str],
                 sheet_of_image: dict, map_seed: int, map_width: int):
        self.ctx = ctx
        self.clutter = list(clutter)
        self.map_width = int(map_width)
        defs, cls, masks = resolve_clutter(self.clutter, floor_class, class_paths, sheet_of_image)
        ctx.upload_clutter(defs, cls, masks, map_seed)
        self.sprites = np.zeros(0, dtype=SPRITE_DTYPE)
        self.spritesByPosition: dict[int, list] = {}

    def Update(self, cam_x: float, cam_y: float, cam_heading: float, fov_degrees: float, clutter_distance: float,
               moved: bool = True, forceUpdate: bool = False):
        if not moved and not forceUpdate:          # clutter.go:26-29
            return
        if not self.clutter or clutter_distance <= 0:  # clutter.go:31-34
            return
        q = clutter_queries([(cam_x, cam_y, cam_heading)], fov_degrees)
        cap = CLUTTER_MAX_ANGLES * (int(clutter_distance) + 1) * len(self.clutter)
        counts, recs, sprites = self.ctx.clutter_batch(q, clutter_distance, cap, want_sprites=True)
        n = int(counts[0])
        self.sprites = sprites[0, :n].copy()
        by_pos: dict[int, list] = {}
        for k in range(n):
            r = recs[0, k]
            pos_id = (int(r["cell_x"]) - 1) * self.map_width + int(r["cell_y"])  # clutter.go:55
            by_pos.setdefault(pos_id, [None] * len(self.clutter))[int(r["def"])] = k  # index into self.sprites
        self.spritesByPosition = by_pos
