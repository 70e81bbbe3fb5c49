"""Host-side mirror of model.Pathing (game/model/pathing.go), in Python because no Go toolchain exists in this
image.  Same names and argument meaning as the Go side:

  initPathing(mission)                         game/model/pathing.go:94-112   (the level-0 grid is the tile world)
  (*Pathing).FindPath(startPos, finishPos)     game/model/pathing.go:123-186

The search and the curve run on the device (pm3d_find_path_batch); FindPath here only packs the query and shapes
the answer the way the Go method returns it: (points, error).  FindPaths is the batched form the device is for —
one query per AI unit of a tick (call site game/ai_movement.go:64).
"""
from __future__ import annotations

import numpy as np

from .context import Context


class Pathing:
    def __init__(self, ctx: Context, max_points: int = 4096):
        """`ctx` holds the mission's grid (pm3d_upload_grid): initPathing reads Level(0) of the same map."""
        self.ctx = ctx
        self.max_points = max_points

    def FindPaths(self, starts, finishes):
        """Batched FindPath: list of (points [k][2] ndarray, error string or None)."""
        q = np.concatenate([np.asarray(starts, dtype=np.float64).reshape(-1, 2),
                            np.asarray(finishes, dtype=np.float64).reshape(-1, 2)], axis=1)
        pts, cnt = self.ctx.find_path_batch(q, self.max_points)
        while len(cnt) and int(cnt.max()) > self.max_points:   # a path longer than the output: ask again, larger
            self.max_points = int(cnt.max())
            pts, cnt = self.ctx.find_path_batch(q, self.max_points)
        out = []
        for i in range(len(q)):
            if cnt[i] < 0:
                out.append((np.zeros((0, 2)), "unable to find path for (%0.0f,%0.0f) -> (%0.0f,%0.0f)" % tuple(q[i])))
            else:
                out.append((pts[i, :cnt[i]].copy(), None))
        return out

    def FindPath(self, startPos, finishPos):
        return self.FindPaths([startPos], [finishPos])[0]
