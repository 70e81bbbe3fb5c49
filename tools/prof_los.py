"""One c5 LOS batch (for ncu): k_los_fast + k_los."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
n = 1 << 24
density = float(os.environ.get("LOS_DENSITY", "0.10"))
ctx = pm3d.Context(0)
d_rays, d_out = ctx.device_alloc(n * 32), ctx.device_alloc(n)
grid = W.synthetic_levels(256, 256, 1, density, 42)
segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
rays = W.los_rays(grid[0], n, 43)
ctx.upload_segments(segs)
ctx.h2d(d_rays, rays)
for _ in range(3):
    ctx._check(ctx.lib.pm3d_los_batch(ctx.h, C.c_void_p(d_rays), n, C.c_void_p(d_out), 7), "los")
ctx.sync()
