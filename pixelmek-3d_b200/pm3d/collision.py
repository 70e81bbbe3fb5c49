"""Host-side mirror of the reference's collision boundary, in Python because
no Go toolchain exists in this image.  Same names and argument meaning as the Go methods:

  (*Game).lineOfSight(source, target)                       game/collision.go:40-65
  (*Game).getValidMove(entity, x, y, z, checkAlternate)     game/collision.go:69-355

Every geometric test runs on the device (pm3d_los_batch / pm3d_valid_move_batch).  What stays here is what the
reference keeps in scalar control flow on top of those tests: the checkAlternate escape heuristics
(collision.go:204-306), which re-issue getValidMove with modified targets and use the entity heading through
atan2/cos/sin (host transcendentals, as everywhere in this package).  Nothing here computes an intersection
on the CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from ._lib import COLLIDER_DTYPE, MOVER_DTYPE
from .context import Context

CLIP_DISTANCE = 0.01  # game/game.go:33
ANCHOR_BOTTOM, ANCHOR_CENTER, ANCHOR_TOP = 0, 1, 2
HALF_PI, PI, PI2 = math.pi / 2, math.pi, math.pi * 2


@dataclass
class Entity:
    """The subset of model.Entity (game/model/entity.go:12-50) getValidMove reads."""
    x: float
    y: float
    z: float = 0.0
    collision_radius: float = 0.0
    collision_height: float = 0.0
    anchor: int = ANCHOR_BOTTOM
    destroyed: bool = False
    heading: float = 0.0
    velocity: float = 0.0
    parent: "Entity | None" = None
    index: int = field(default=-1, compare=False)  # position in the collider list, set by Game.set_colliders

    def Pos(self):
        return self.x, self.y

    def PosZ(self):
        return self.z


@dataclass
class EntityCollision:
    """game/collision.go:14-18"""
    entity: Entity
    collision: tuple
    collisionZ: float


def angle_distance(a: float, b: float) -> float:
    """model.AngleDistance (game/model/math.go:149-152)"""
    return math.fmod(math.fmod(b - a + PI, PI2) - PI2, PI2) + PI


def nearly_equal(a: float, b: float, epsilon: float) -> bool:
    """geom.NearlyEqual"""
    if a == b:
        return True
    diff = abs(a - b)
    tiny = 5e-324
    if a == 0.0 or b == 0.0 or diff < tiny:
        return diff < epsilon * tiny
    return diff / (abs(a) + abs(b)) < epsilon


class Game:
    """The collision-facing part of game.Game: g.collisionMap, g.player.Unit, g.sprites (collision types only)."""

    def __init__(self, ctx: Context, collision_map: np.ndarray):
        self.ctx = ctx
        self.collisionMap = np.ascontiguousarray(collision_map, dtype=np.float64).reshape(-1, 4)
        ctx.upload_segments(self.collisionMap)
        self.player: Entity | None = None
        self.colliders: list[Entity] = []
        self.set_colliders(None, [])

    def set_colliders(self, player: Entity | None, sprites: list[Entity]):
        """Snapshot g.player.Unit and the sprites of the collision types (collision.go:20-29) for this tick."""
        self.player = player
        self.colliders = ([player] if player is not None else []) + list(sprites)
        arr = np.zeros(len(self.colliders), dtype=COLLIDER_DTYPE)
        for i, e in enumerate(self.colliders):
            e.index = i
            arr[i] = (e.x, e.y, e.z, e.collision_radius, e.collision_height, e.anchor, int(e.destroyed), [0] * 6)
        self.ctx.upload_colliders(arr, 0 if player is not None else -1)

    # ------------------------------------------------------------------ lineOfSight
    def lineOfSight(self, source: Entity | None, target: Entity | None) -> bool:
        if source is None or target is None:
            return False
        return bool(self.ctx.los_batch(np.array([[source.x, source.y, target.x, target.y]]))[0])

    # ------------------------------------------------------------------ getValidMove
    def _mover(self, e: Entity, mx: float, my: float, mz: float):
        m = np.zeros(1, dtype=MOVER_DTYPE)
        pa = e.parent
        parent = pa.index if pa is not None and 0 <= pa.index < len(self.colliders) and self.colliders[pa.index] is pa else -1
        self_i = e.index if 0 <= e.index < len(self.colliders) and self.colliders[e.index] is e else -1
        m[0] = (e.x, e.y, e.z, mx, my, mz, e.collision_radius, e.collision_height, self_i, parent, e.anchor,
                int(e.destroyed), [0] * 6)
        return m

    def movers(self, moves) -> np.ndarray:
        """[(entity, moveX, moveY, moveZ), ...] -> pm3d_mover records for one batched call."""
        return np.concatenate([self._mover(e, x, y, z) for e, x, y, z in moves]) if moves else np.zeros(0, MOVER_DTYPE)

    def getValidMoveBatch(self, moves):
        """checkAlternate=false for many entities in ONE device call (the projectile fan-out of
        game/combat.go:206-239).  Returns the raw pm3d_move_result records."""
        return self.ctx.valid_move_batch(self.movers(moves), CLIP_DISTANCE)

    def getValidMove(self, entity: Entity, moveX: float, moveY: float, moveZ: float, checkAlternate: bool):
        """Returns ((x, y), z, isCollision, collisions) like the Go method; `collisions` holds collisions[0]
        (the only element the callers read) when there is one."""
        r = self.ctx.valid_move_batch(self._mover(entity, moveX, moveY, moveZ), CLIP_DISTANCE)[0]
        collisions = []
        if r["n_collisions"] > 0:
            collisions.append(EntityCollision(self.colliders[int(r["first_entity"])],
                                              (float(r["first_x"]), float(r["first_y"])), float(r["first_z"])))
        is_collision = bool(r["collided"])
        pos, z = (float(r["x"]), float(r["y"])), float(r["z"])
        if not (r["n_points"] > 0 and checkAlternate):
            return pos, z, is_collision, collisions
        posX, posY, posZ = entity.x, entity.y, entity.z
        if abs(moveZ - posZ) > 0:                                                     # collision.go:205-217
            zP, zZ, zCollide, _ = self.getValidMove(entity, posX, posY, moveZ, False)
            if not zCollide:
                return zP, zZ, is_collision, collisions
            zP, zZ, _, _ = self.getValidMove(entity, moveX, moveY, posZ, True)
            return zP, zZ, is_collision, collisions
        if collisions:                                                                # :219-249
            # the device returns intersectPoints[minI] (nearest to the entity, first on ties)
            nx, ny = float(r["near_x"]), float(r["near_y"])
            if math.isnan(nx):  # every intersection point is infinitely far (the reference would panic at [-1])
                return pos, z, is_collision, collisions
            intersectAngle = math.atan2(ny - posY, nx - posX)
            dx, dy, dz = moveX - posX, moveY - posY, moveZ - posZ
            checkDist = math.sqrt(dx * dx + dy * dy + dz * dz)
            hi = entity.collision_radius - 0.001
            escapeDist = 0.0 if checkDist < 0 else (hi if checkDist > hi else checkDist)
            if angle_distance(entity.heading, intersectAngle) >= 0:
                escapeAngle = intersectAngle - HALF_PI
            else:
                escapeAngle = intersectAngle + HALF_PI
            escapeX, escapeY = posX + escapeDist * math.cos(escapeAngle), posY + escapeDist * math.sin(escapeAngle)
            nP, nZ, _, _ = self.getValidMove(entity, escapeX, escapeY, moveZ, False)
            return nP, nZ, is_collision, collisions
        identityX, identityY = abs(0.0 + 1.0 * math.cos(entity.heading)), abs(0.0 + 1.0 * math.sin(entity.heading))
        if identityX >= identityY:                                                    # :255-279
            nP, nZ, xCollide, _ = self.getValidMove(entity, moveX, posY, moveZ, False)
            if not xCollide:
                return nP, nZ, is_collision, collisions
            if nearly_equal(posY, moveY, 0.001):
                dY = entity.velocity / 2
                if moveY - posY == 0:
                    if math.copysign(1.0, moveX) < 0:
                        dY *= -1
                elif math.copysign(1.0, moveY - posY) < 0:
                    dY *= -1
                moveY = posY + dY
            nP, nZ, _, _ = self.getValidMove(entity, posX, moveY, moveZ, False)
            return nP, nZ, is_collision, collisions
        nP, nZ, yCollide, _ = self.getValidMove(entity, posX, moveY, moveZ, False)     # :280-304
        if not yCollide:
            return nP, nZ, is_collision, collisions
        if nearly_equal(posX, moveX, 0.001):
            dX = entity.velocity / 2
            if moveX - posX == 0:
                if math.copysign(1.0, moveY) < 0:
                    dX *= -1
            elif math.copysign(1.0, moveX - posX) < 0:
                dX *= -1
            moveX = posX + dX
        nP, nZ, _, _ = self.getValidMove(entity, moveX, posY, moveZ, False)
        return nP, nZ, is_collision, collisions
