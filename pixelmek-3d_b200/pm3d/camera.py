"""Host-side mirror of the reference's camera boundary, in Python because no Go
toolchain exists in this image.  Same names, argument meaning and error behaviour as the Go
interfaces the game drives, so the parity tests read like tests of raycaster.Camera:

  raycaster.NewCamera(w, h, texSize, Map, TextureHandler)   game/mission.go:74
  Camera.Update(sprites) / GetConvergenceSprite() / Draw()   game/scene_game.go:161-167
  setters/getters                                            game/camera.go:31-52, game/config.go:248-291,
                                                             game/input.go:764-792, game/mission.go:75-92

The callbacks (Map.Level, TextureHandler.TextureAt/FloorTextureAt, Sprite.*) are resolved
BEFORE the device call: once per mission for the grid and textures (flatten_scene), once per
frame for the sprite snapshot.  Transcendentals are evaluated here, on the host; the device
path receives dir/plane/pitch already evaluated (include/pm3d.h).
"""
from __future__ import annotations

import math

import numpy as np

from ._lib import SPRITE_DTYPE, VIEW_DTYPE
from .context import Context
from .scene import SceneData

ANCHOR_BOTTOM, ANCHOR_CENTER, ANCHOR_TOP = 0, 1, 2  # raycaster.AnchorBottom/Center/Top


# --------------------------------------------------------------------------- callback interfaces


class LevelMap:
    """raycaster.Map as model.Map implements it (game/model/map.go:245-264)."""

    def __init__(self, levels: np.ndarray, num_raycast_levels: int | None = None):
        self.levels = np.ascontiguousarray(levels, dtype=np.uint16)  # [level][x][y]
        self.num_raycast_levels = int(num_raycast_levels or self.levels.shape[0])

    def Size(self):
        if self.levels.size == 0:
            return 0, 0
        return int(self.levels.shape[1]), int(self.levels.shape[2])

    def NumLevels(self) -> int:
        return self.num_raycast_levels

    def Level(self, level_num: int) -> np.ndarray:
        n = self.levels.shape[0]
        return self.levels[level_num] if level_num < n else self.levels[n - 1]

    def IsWallAt(self, level_num: int, x: int, y: int) -> bool:
        return bool(self.Level(level_num)[x][y] > 0)


class TextureHandler:
    """raycaster.TextureHandler as game/texture/texture.go implements it.

    textures: {texIndex: (image, sideX, sideY)} with image objects (uint8 [T][T][4] premultiplied
    arrays) or None — MapTexture + GetImage (game/model/map.go:100-115)."""

    def __init__(self, level_map: LevelMap, textures: dict, floor_default, floor_grid=None, floor_images=None):
        self.map = level_map
        self.textures = textures
        self.floor_default = floor_default
        self.floor_grid = floor_grid      # int [x][y] index into floor_images, or None
        self.floor_images = floor_images  # list of images; index 0 = default
        self.render_floor_tex = True

    def SetRenderFloorTex(self, on: bool):
        self.render_floor_tex = bool(on)

    def TextureAt(self, x: int, y: int, level_num: int, side: int):
        level = self.map.Level(level_num)
        if level is None or level.shape[0] == 0 or level.shape[1] == 0:
            return None
        tex_num = -1
        if 0 <= x < level.shape[0] and 0 <= y < level.shape[1]:
            tex_num = int(level[x][y])
        if tex_num <= 0:
            return None
        entry = self.textures.get(tex_num)
        if entry is None:
            return None
        image, side_x, side_y = entry
        if side == 0 and side_y is not None:
            return side_y
        if side == 1 and side_x is not None:
            return side_x
        return image

    def FloorTextureAt(self, x: int, y: int):
        if x < 0 or y < 0:
            return None
        if not self.render_floor_tex:
            return None
        if self.floor_grid is not None:
            idx = int(self.floor_grid[x][y])  # IndexError past the edge, like the Go panic
            if idx > 0:
                return self.floor_images[idx]
        return self.floor_default


class Sprite:
    """raycaster.Sprite as game/render/sprites/sprite.go:49-91 implements it."""

    def __init__(self, x, y, z=0.0, scale=1.0, anchor=ANCHOR_BOTTOM, texture=None, texture_rect=None,
                 illumination=0.0, focusable=True):
        self.x, self.y, self.z = float(x), float(y), float(z)
        self.scale = float(scale)
        self.anchor = int(anchor)
        self.texture = texture
        self.texture_rect = texture_rect if texture_rect is not None else (0, 0, texture.shape[1], texture.shape[0])
        self.illumination = float(illumination)
        self.focusable = bool(focusable)
        self.screen_rect = None

    def Pos(self): return (self.x, self.y)
    def PosZ(self): return self.z
    def Scale(self): return self.scale
    def VerticalAnchor(self): return self.anchor
    def Texture(self): return self.texture
    def TextureRect(self): return self.texture_rect
    def Illumination(self): return self.illumination
    def SetScreenRect(self, rect): self.screen_rect = rect
    def IsFocusable(self): return self.focusable

    def ScreenRect(self, render_scale: float = 1.0):
        """sprite.go:300-313"""
        if self.screen_rect is None:
            return None
        if render_scale == 1:
            return self.screen_rect
        x0, y0, x1, y1 = self.screen_rect
        return (int(x0 * 1 / render_scale), int(y0 * 1 / render_scale), int(x1 * 1 / render_scale), int(y1 * 1 / render_scale))


def sheet_slices(img_w: int, img_h: int, columns: int, rows: int):
    """GetSpriteSheetSlices rectangles (sprite.go:161-185), including the reference's
    one-pixel-short cells: image.Rect(x, y, x+spriteW-1, y+spriteH-1)."""
    w_f, h_f = img_w / columns, img_h / rows
    sw, sh = int(w_f), int(h_f)
    rects = []
    for r in range(rows):
        y = int(r * h_f)
        for c in range(columns):
            x = int(c * w_f)
            rects.append((x, y, x + sw - 1, y + sh - 1))
    return rects


# --------------------------------------------------------------------------- flattening


def flatten_scene(level_map: LevelMap, tex: TextureHandler, tex_size: int, sky=None, floor_box=None,
                  sheets: list | None = None) -> SceneData:
    """Resolve the Map / TextureHandler callbacks into the arrays the C-ABI uploads.

    TextureAt is called once per distinct (level, texture index, side) on a representative
    cell; FloorTextureAt once per cell of the floor grid's distinct entries."""
    w, h = level_map.Size()
    stored = level_map.levels.shape[0]
    images: list = []
    ids: dict = {}

    def tile_of(img):
        if img is None:
            return -1
        key = id(img)
        if key not in ids:
            a = np.asarray(img)
            if a.shape != (tex_size, tex_size, 4):
                raise ValueError(f"texture must be {tex_size}x{tex_size} RGBA, got {a.shape}")
            ids[key] = len(images)
            images.append(a)
        return ids[key]

    lut: dict = {}
    for lvl in range(min(stored, level_map.NumLevels())):
        grid = level_map.Level(lvl)
        vals, first = np.unique(grid, return_index=True)
        for v, fi in zip(vals, first):
            v = int(v)
            if v <= 0:
                continue
            x, y = divmod(int(fi), h)
            t0 = tile_of(tex.TextureAt(x, y, lvl, 0))
            t1 = tile_of(tex.TextureAt(x, y, lvl, 1))
            prev = lut.get(v)
            if prev is not None and prev != (t0, t1):
                raise ValueError(f"texture index {v} resolves differently on level {lvl}")
            lut[v] = (t0, t1)

    floor_tiles = np.full((w, h), -1, dtype=np.int16)
    if tex.render_floor_tex:
        if tex.floor_grid is not None:
            fg = np.asarray(tex.floor_grid)
            for v in np.unique(fg):
                xs, ys = np.nonzero(fg == v)
                floor_tiles[fg == v] = tile_of(tex.FloorTextureAt(int(xs[0]), int(ys[0])))
        else:
            floor_tiles[:, :] = tile_of(tex.FloorTextureAt(0, 0))

    sky_tile, floorbox_tile = tile_of(sky), tile_of(floor_box)
    keys = sorted(lut)
    return SceneData(
        grid=level_map.levels,
        num_levels=level_map.NumLevels(),
        wall_tex_index=np.array(keys, dtype=np.int32),
        wall_tile_side0=np.array([lut[k][0] for k in keys], dtype=np.int32),
        wall_tile_side1=np.array([lut[k][1] for k in keys], dtype=np.int32),
        floor_tiles=floor_tiles,
        atlas=np.stack(images).astype(np.uint8) if images else np.zeros((1, tex_size, tex_size, 4), np.uint8),
        sky_tile=sky_tile,
        floorbox_tile=floorbox_tile,
        sheets=list(sheets or []),
    )


# --------------------------------------------------------------------------- camera state -> pm3d_view


class CameraState:
    """The scalar state behind the K5 setters/getters, and its lowering to pm3d_view."""

    def __init__(self, w: int, h: int):
        self.w, self.h = int(w), int(h)
        self.pos = (1.0, 1.0)
        self.pos_z = 0.5
        self.heading = 0.0
        self.pitch_angle = 0.0
        self.fov_angle = math.radians(60.0)
        self.fov_depth = 1.0
        self.render_distance = -1.0
        self.light_falloff = -100.0
        self.global_illumination = 300.0
        self.min_light = (0, 0, 0)
        self.max_light = (255, 255, 255)

    def lowered(self) -> dict:
        """dir = fovDepth*(cos, sin)(heading); plane = perp(dir)*tan(fov/2);
        pitch = clamp(int(h*fovDepth*tan(pitchAngle)), -h/2, int(h*fovDepth)); camZ = (posZ-0.5)*h.
        [R] raycaster-go Camera.SetHeadingAngle/SetFovAngle/SetPitchAngle/SetPositionZ."""
        dx = self.fov_depth * math.cos(self.heading)
        dy = self.fov_depth * math.sin(self.heading)
        t = math.tan(self.fov_angle / 2.0)
        pitch = int(self.h * self.fov_depth * math.tan(self.pitch_angle))
        lo, hi = -(self.h // 2), int(self.h * self.fov_depth)
        pitch = lo if pitch < lo else (hi if pitch > hi else pitch)
        return dict(
            pos_x=self.pos[0], pos_y=self.pos[1], cam_z_px=(self.pos_z - 0.5) * self.h,
            dir_x=dx, dir_y=dy, plane_x=dy * t, plane_y=-dx * t, pitch_px=pitch, w=self.w, h=self.h,
            render_distance=self.render_distance, light_falloff=self.light_falloff,
            global_illumination=self.global_illumination, min_light=tuple(self.min_light), max_light=tuple(self.max_light),
        )

    def view_record(self) -> np.ndarray:
        p = self.lowered()
        v = np.zeros(1, dtype=VIEW_DTYPE)
        for k in ("pos_x", "pos_y", "cam_z_px", "dir_x", "dir_y", "plane_x", "plane_y", "pitch_px", "render_distance",
                  "light_falloff", "global_illumination"):
            v[0][k] = p[k]
        v[0]["min_light"][:3] = p["min_light"][:3]
        v[0]["max_light"][:3] = p["max_light"][:3]
        v[0]["min_light"][3] = v[0]["max_light"][3] = 255
        return v


def snapshot_sprites(sprites, sheet_ids: dict) -> np.ndarray:
    """Per-frame SoA snapshot of []raycaster.Sprite (the accessor calls of sprite.go:49-91)."""
    out = np.zeros(len(sprites), dtype=SPRITE_DTYPE)
    for i, s in enumerate(sprites):
        px, py = s.Pos()
        rect = s.TextureRect()
        tex = s.Texture()
        out[i]["x"], out[i]["y"], out[i]["z"] = px, py, s.PosZ()
        out[i]["scale"] = s.Scale()
        out[i]["illumination"] = s.Illumination()
        out[i]["sheet"] = sheet_ids.get(id(tex), -1) if tex is not None else -1
        out[i]["rx0"], out[i]["ry0"], out[i]["rx1"], out[i]["ry1"] = rect
        out[i]["anchor"] = s.VerticalAnchor()
        out[i]["focusable"] = 1 if s.IsFocusable() else 0
    return out


class Camera:
    """Drop-in for *raycaster.Camera on the calls the game makes."""

    def __init__(self, w: int, h: int, tex_size: int, level_map: LevelMap, tex: TextureHandler,
                 ctx: Context | None = None, sheets: list | None = None, max_sprites: int = 2048):
        # NewCamera: game/mission.go:74
        self.ctx = ctx or Context(0)
        self.tex_size = int(tex_size)
        self.level_map, self.tex = level_map, tex
        self.state = CameraState(w, h)
        self.sky = self.floor_box = None
        self.sheets = list(sheets or [])
        self.max_sprites = max_sprites
        self.always_set_sprite_screen_rect = False
        self._dirty_scene = True
        self._sprites = []
        self._last = None

    # -- mission-time
    def _ensure_scene(self):
        if not self._dirty_scene:
            return
        scene = flatten_scene(self.level_map, self.tex, self.tex_size, self.sky, self.floor_box, self.sheets)
        self.ctx.upload_scene(scene)
        self.ctx.set_view(self.state.w, self.state.h, 1, self.max_sprites)
        self._sheet_ids = {id(s): i for i, s in enumerate(self.sheets)}
        self.scene = scene
        self._dirty_scene = False

    def SetFloorTexture(self, img):  # game/mission.go:79
        self.floor_box = img
        self._dirty_scene = True

    def SetSkyTexture(self, img):  # game/mission.go:82
        self.sky = img
        self._dirty_scene = True

    def SetViewSize(self, w: int, h: int):  # game/config.go:249
        self.state.w, self.state.h = int(w), int(h)
        if not self._dirty_scene:
            self.ctx.set_view(self.state.w, self.state.h, 1, self.max_sprites)

    def ViewSize(self):
        return self.state.w, self.state.h

    # -- K5 setters / getters
    def SetPosition(self, pos): self.state.pos = (float(pos[0]), float(pos[1]))          # game/camera.go:31
    def SetPositionZ(self, z): self.state.pos_z = float(z)                               # game/camera.go:32
    def SetHeadingAngle(self, a): self.state.heading = float(a)                          # game/camera.go:33
    def SetPitchAngle(self, a): self.state.pitch_angle = float(a)                        # game/camera.go:34
    def GetPosition(self): return self.state.pos                                         # game/camera.go:39
    def GetPositionZ(self): return self.state.pos_z                                      # game/camera.go:48
    def SetRenderDistance(self, d): self.state.render_distance = float(d)                # game/mission.go:75
    def SetAlwaysSetSpriteScreenRect(self, on): self.always_set_sprite_screen_rect = bool(on)  # game/mission.go:76
    def SetLightFalloff(self, f): self.state.light_falloff = float(f)                    # game/mission.go:90
    def SetGlobalIllumination(self, g): self.state.global_illumination = float(g)        # game/mission.go:91

    def SetLightRGB(self, min_rgb, max_rgb):                                             # game/mission.go:92
        self.state.min_light, self.state.max_light = tuple(min_rgb[:3]), tuple(max_rgb[:3])

    def SetFovAngle(self, fov_degrees: float, fov_depth: float):                         # game/config.go:290
        self.state.fov_angle = math.radians(fov_degrees)
        self.state.fov_depth = float(fov_depth)

    def FovDepth(self): return self.state.fov_depth                                      # game/mission.go:97
    def FovRadians(self): return self.state.fov_angle                                    # game/hud.go:844
    def FovRadiansVertical(self):                                                        # game/hud.go:878
        return 2.0 * math.atan(math.tan(self.state.fov_angle / 2.0) * (self.state.h / self.state.w))

    # -- K2 / K3 / K4
    def Update(self, sprites):  # game/scene_game.go:161
        self._ensure_scene()
        self._sprites = list(sprites)
        snap = snapshot_sprites(self._sprites, self._sheet_ids)
        res = self.ctx.render(self.state.view_record(), snap)
        self._last = res
        for s, rc in zip(self._sprites, res["rects"]):
            s.SetScreenRect(None if rc[0] > rc[2] else (int(rc[0]), int(rc[1]), int(rc[2]), int(rc[3])))

    def GetConvergenceSprite(self):  # game/scene_game.go:164
        if self._last is None:
            return None
        i = int(self._last["conv"][0]["sprite"])
        return self._sprites[i] if i >= 0 else None

    def GetConvergenceDistance(self):
        return float(self._last["conv"][0]["distance"]) if self._last is not None else -1.0

    def GetConvergencePoint(self):
        c = self._last["conv"][0]
        return float(c["point_x"]), float(c["point_y"])

    def Draw(self, screen: np.ndarray | None = None) -> np.ndarray:  # game/scene_game.go:167
        """Returns (or fills) the [h][w][4] premultiplied RGBA8 frame of the last Update."""
        if self._last is None:
            raise RuntimeError("Draw before Update")
        frame = self._last["rgba"][0]
        if screen is not None:
            screen[...] = frame
            return screen
        return frame
