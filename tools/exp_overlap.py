"""c2 resident frame rate for PM3D_OPT_OVERLAP = 0, 1, 2 (tools only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
c2 = W.config_c2(n_poses=24)
views, spr, off = c2["views"], c2["sprites"], c2["offsets"]
ctx.upload_scene(c2["scene"]); ctx.set_view(1920, 1080, 1, 200)
FB = 1920 * 1080 * 4; ring = 24
dv = ctx.device_alloc(views.nbytes); ds = ctx.device_alloc(spr.nbytes); do = ctx.device_alloc(8); dout = ctx.device_alloc(ring * FB)
ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, 200], np.int32))
def frame(i):
    r = i % ring
    ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, 200, dout + r * FB)
for ov in (0, 1, 2):
    ctx.set_option(pm3d._lib.OPT_OVERLAP, ov)
    for i in range(50): frame(i)
    ctx.sync(); reps = 2000
    ctx.timer_start(0)
    for i in range(reps): frame(i)
    ctx.timer_stop(0); ms = ctx.timer_ms(0) / reps
    print(f"overlap={ov}: {ms*1e3:.1f} us/frame = {1e3/ms:.0f} frames/s")
