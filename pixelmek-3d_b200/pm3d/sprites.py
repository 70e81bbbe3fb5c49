"""Host-side mirror of (*Game).getRaycastSprites (game/sprites.go:102-145; SURVEY.md §8a row SL), in Python because no
Go toolchain exists in this image.  The proximity filter over the map's sprites and the concatenation with clutter and
extras run on the device (pm3d_raycast_sprites_batch); this module only shapes the arguments."""
from __future__ import annotations

import numpy as np

from ._lib import LIST_QUERY_DTYPE, SPRITE_DTYPE
from .context import Context


def list_queries(cameras, targets=None) -> np.ndarray:
    """pm3d_list_query records for cameras = iterable of (camX, camY); targets = index of g.player.Target() in the
    world list per camera (-1 = none)."""
    cams = list(cameras)
    q = np.zeros(len(cams), dtype=LIST_QUERY_DTYPE)
    for i, c in enumerate(cams):
        q["cam_x"][i], q["cam_y"][i] = c[0], c[1]
    q["target"] = -1 if targets is None else np.asarray(targets, dtype=np.int32)
    return q


def getRaycastSprites(ctx: Context, world: np.ndarray, cam_x: float, cam_y: float, render_distance: float,
                      target: int = -1, clutter: np.ndarray | None = None, extras: np.ndarray | None = None):
    """One camera's sprite list for Camera.Update: (sprites, index).  clutter = ClutterHandler.sprites; extras = the
    reticle lead / current nav point / player sprite records the game appends (sprites.go:127-142)."""
    cl = None
    if clutter is not None:
        clutter = np.ascontiguousarray(clutter, dtype=SPRITE_DTYPE).reshape(1, -1)
        if clutter.shape[1]:
            cl = (np.array([clutter.shape[1]], dtype=np.int32), clutter)
    ex = [np.ascontiguousarray(extras, dtype=SPRITE_DTYPE).reshape(-1)] if extras is not None else None
    sprites, _, index = ctx.raycast_sprites_batch(world, list_queries([(cam_x, cam_y)], [target]), render_distance, cl, ex)
    return sprites, index
