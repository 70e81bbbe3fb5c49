"""What one rank of an 8-GPU c4 / c5 sweep does, on ONE GPU (1-rank communicator): n views (default 512 = 4096 / 8) through
pm3d_comm_render_views and n rays (default 2 Mi = 16 Mi / 8) through pm3d_comm_los_batch, back to back, per-batch time and
per-kernel times.  Tools only:  python tools/exp_shard.py [n_views] [n_rays] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np
import pm3d
from pm3d import workloads as W, generate_wall_collision_lines

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = pm3d.Context(0)
ctx.comm_init(pm3d.comm_unique_id(), 0, 1)
cfg = W.config_c4(n_views=nv)
ctx.upload_scene(cfg["scene"]); ctx.set_view(320, 200, nv, len(cfg["sprites"]))
v, s, o = cfg["views"], cfg["sprites"], cfg["offsets"]
dv, ds, do = ctx.device_alloc(v.nbytes), ctx.device_alloc(s.nbytes), ctx.device_alloc(o.nbytes)
dout = ctx.device_alloc(nv * 320 * 200 * 4)
dconv, dsum = ctx.device_alloc(nv * 32), ctx.device_alloc(nv * 8)
ctx.h2d(dv, v); ctx.h2d(ds, s); ctx.h2d(do, o)
call = lambda: ctx.comm_render_views(dv, nv, ds, do, len(s), dout, nv, dconv, dsum)
for _ in range(3): call()
ctx.sync()
ctx.ktimer_begin(reps); ctx.timer_start(0)
for _ in range(reps): call()
ctx.timer_stop(0); ms = ctx.timer_ms(0) / reps
kt = ctx.ktimer_collect(reps)
print(f"c4 shard, {nv} views: {ms*1e3:.1f} us per batch (front {kt[:,0].mean()*1e3:.1f}, compose {kt[:,2].mean()*1e3:.1f}) = {nv/ms/1e3:.2f} M views/s per GPU")
grid = W.synthetic_levels(256, 256, 1, 0.10, 42)
segs = generate_wall_collision_lines(grid[0], True, W.CLIP_DISTANCE)
rays = W.los_rays(grid[0], nr, 43)
ctx.upload_segments(segs)
dr, dlo = ctx.device_alloc(rays.nbytes), ctx.device_alloc(nr)
ctx.h2d(dr, rays)
call = lambda: ctx.comm_los_batch(dr, nr, nr, dlo)
for _ in range(3): call()
ctx.sync()
ctx.timer_start(1)
for _ in range(reps): call()
ctx.timer_stop(1); ms = ctx.timer_ms(1) / reps
print(f"c5 shard, {nr} rays: {ms*1e3:.1f} us per batch = {nr/ms/1e6:.2f} G rays/s per GPU")
