"""Clutter batch run for ncu: ClutterHandler.Update for N cameras on the bundled map, resident."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import ctypes as C
import numpy as np
import pm3d
from pm3d import clutter as cl, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
scene, g = W.proving_grounds_scene()
clutter = [cl.MapClutter(str(im), float(f), float(h), str(m) or None) for im, m, f, h in
           zip(g["clutter_image"], g["clutter_match"], g["clutter_frequency"], g["clutter_height"])]
defs, cls, masks = cl.resolve_clutter(clutter, g["floor_grid"], [str(x) for x in g["floor_images"]],
                                      {c.image: (3, int(w), int(h)) for c, (w, h) in zip(clutter, g["clutter_image_size"])})
ctx = pm3d.Context(0)
ctx.upload_scene(scene); ctx.upload_clutter(defs, cls, masks, int(g["map_seed"]))
q = cl.clutter_queries([(p[0], p[1], p[3]) for p in W.free_cell_poses(scene, n, 4096)], 70.0)
cap = 21 * 26 * len(clutter)
d_q, d_cnt, d_rec, d_spr = ctx.device_alloc(q.nbytes), ctx.device_alloc(4 * n), ctx.device_alloc(32 * n * cap), ctx.device_alloc(64 * n * cap)
ctx.h2d(d_q, q)
call = lambda: ctx._check(ctx.lib.pm3d_clutter_batch(ctx.h, C.c_void_p(d_q), n, 25.0, cap, C.c_void_p(d_cnt), C.c_void_p(d_rec), C.c_void_p(d_spr), 7), "clutter")
call(); ctx.sync()
ctx.timer_start(0)
for _ in range(reps): call()
ctx.timer_stop(0)
counts = np.zeros(n, np.int32); ctx.d2h(counts, d_cnt)
print(f"{n} cameras: {ctx.timer_ms(0) / reps * 1e3:.1f} us/batch, {int(counts.sum())} sprites")
