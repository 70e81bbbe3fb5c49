import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
n = 1 << 24
grid = W.synthetic_levels(256, 256, 1, 0.10, 42)
segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
rays = W.los_rays(grid[0], n, 43)
ctx = pm3d.Context(0)
ctx.upload_segments(segs)
d_rays, d_out = ctx.device_alloc(rays.nbytes), ctx.device_alloc(n)
ctx.h2d(d_rays, rays)
for _ in range(2):
    ctx._check(ctx.lib.pm3d_los_batch(ctx.h, C.c_void_p(d_rays), n, C.c_void_p(d_out), 7), "los")
ctx.sync()
ctx.timer_start(0)
for _ in range(3):
    ctx._check(ctx.lib.pm3d_los_batch(ctx.h, C.c_void_p(d_rays), n, C.c_void_p(d_out), 7), "los")
ctx.timer_stop(0)
ms = ctx.timer_ms(0) / 3
print(f"waves={os.environ.get('PM3D_LOS_WAVES','1')}: {ms:.3f} ms, {n/ms/1e6:.2f} Grays/s, {33*n/ms/1e6:.1f} GB/s algorithmic")
