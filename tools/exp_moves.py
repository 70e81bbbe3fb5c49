"""Timing of the full getValidMove batch (tools only)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
n = 1 << 22
grid = W.synthetic_levels(256, 256, 1, 0.10, 42)
segs = pm3d.generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
ctx = pm3d.Context(0)
ctx.upload_scene(W.scene_from_grid(grid, 1)); ctx.upload_segments(segs)
for ncol in (0, 64, 512):
    colliders, movers = W.move_workload(grid[0], n, ncol, 77)
    ctx.upload_colliders(colliders, 0 if ncol else -1)
    d_in, d_out = ctx.device_alloc(movers.nbytes), ctx.device_alloc(n * 80)
    ctx.h2d(d_in, movers)
    call = lambda: ctx._check(ctx.lib.pm3d_valid_move_batch(ctx.h, C.c_void_p(d_in), n, W.CLIP_DISTANCE, C.c_void_p(d_out), 7), "move")
    call(); ctx.sync()
    ctx.timer_start(0)
    for _ in range(3): call()
    ctx.timer_stop(0); ms = ctx.timer_ms(0) / 3
    print(f"{ncol:4d} colliders: {ms:.3f} ms, {n/ms/1e3:.0f} M moves/s")
    ctx.device_free(d_in); ctx.device_free(d_out)
