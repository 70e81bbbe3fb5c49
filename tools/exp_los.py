"""LOS batch timing: c5 (256x256, 10 % walls) and sparser maps, two-pass (k_los_fast + k_los) against the exact pass alone."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
n = 1 << 24
ctx = pm3d.Context(0)
d_rays, d_out = ctx.device_alloc(n * 32), ctx.device_alloc(n)
for size, density in ((256, 0.10), (256, 0.02), (256, 0.005), (512, 0.10)):
    grid = W.synthetic_levels(size, size, 1, density, 42)
    segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
    rays = W.los_rays(grid[0], n, 43)
    ctx.upload_segments(segs)
    ctx.h2d(d_rays, rays)
    outs = {}
    for fast in ("1", "0"):
        os.environ["PM3D_LOS_FAST"] = fast
        for _ in range(2):
            ctx._check(ctx.lib.pm3d_los_batch(ctx.h, C.c_void_p(d_rays), n, C.c_void_p(d_out), 7), "los")
        ctx.sync()
        ctx.timer_start(0)
        for _ in range(3):
            ctx._check(ctx.lib.pm3d_los_batch(ctx.h, C.c_void_p(d_rays), n, C.c_void_p(d_out), 7), "los")
        ctx.timer_stop(0)
        ms = ctx.timer_ms(0) / 3
        out = np.zeros(n, dtype=np.uint8); ctx.d2h(out, d_out); outs[fast] = out
        print(f"{size}x{size} {density:.3f} fast={fast}: {ms:.3f} ms, {n/ms/1e6:.2f} Grays/s, {33*n/ms/1e6:.1f} GB/s algorithmic, visible {out.mean():.4f}", flush=True)
    assert np.array_equal(outs["0"], outs["1"]), "two-pass and exact pass disagree"
