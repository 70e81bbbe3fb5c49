"""Flattened mission data: what the Go shim resolves from the game's callback interfaces once
per mission and uploads through the C-ABI.

  raycaster.Map            NumLevels()/Level(l)         game/model/map.go:253-264
  raycaster.TextureHandler TextureAt / FloorTextureAt   game/texture/texture.go:69-97,119-133
  raycaster.Sprite         Texture()/TextureRect()      game/render/sprites/sprite.go:73-79
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def premultiply(straight_rgba: np.ndarray) -> np.ndarray:
    """Straight-alpha RGBA8 -> premultiplied RGBA8 exactly as Go's image/color converts
    color.NRGBA to color.RGBA (what ebiten.NewImageFromImage and image.RGBA hold):
    r16 = r*0x101 * a*0x101 / 0xffff, then >> 8."""
    px = np.ascontiguousarray(straight_rgba, dtype=np.uint8)
    a = px[..., 3].astype(np.uint64) * 0x101
    out = px.copy()
    for ch in range(3):
        c = px[..., ch].astype(np.uint64) * 0x101
        out[..., ch] = ((c * a // 0xFFFF) >> 8).astype(np.uint8)
    return out


@dataclass
class SceneData:
    grid: np.ndarray                 # uint16 [stored_levels][map_w][map_h], index order of Map.Levels
    num_levels: int                  # Map.NumLevels()
    wall_tex_index: np.ndarray       # int32 [n]: texture indices present in the grid
    wall_tile_side0: np.ndarray      # int32 [n]: atlas tile TextureAt(.., side=0) resolves to, -1 = nil
    wall_tile_side1: np.ndarray      # int32 [n]
    floor_tiles: np.ndarray          # int16 [map_w][map_h]: FloorTextureAt tile, -1 = nil
    atlas: np.ndarray                # uint8 [n_tiles][T][T][4] premultiplied
    sky_tile: int = -1
    floorbox_tile: int = -1
    sheets: list = field(default_factory=list)  # uint8 [h][w][4] premultiplied sprite images
    # map constants that ride along for convenience (not uploaded)
    light_falloff: float = -100.0
    global_illumination: float = 300.0
    min_light: tuple = (0, 0, 0)
    max_light: tuple = (255, 255, 255)
    has_boundary_wall: bool = False

    @property
    def map_size(self):
        return int(self.grid.shape[1]), int(self.grid.shape[2])

    @property
    def tex_size(self):
        return int(self.atlas.shape[1])
