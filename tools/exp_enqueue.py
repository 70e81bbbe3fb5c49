"""Host cost of enqueueing resident frames (Python + library + CUDA runtime) against the device time, 1 and 4 lanes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
RING = 24
cfg = W.config_c2(1920, 1080, 200, n_poses=RING)
ctx.upload_scene(cfg["scene"]); ctx.set_view(1920, 1080, 1, 200)
views, spr, off = cfg["views"], cfg["sprites"], cfg["offsets"]
FB = 1920 * 1080 * 4
dv, ds, do, dout = ctx.device_alloc(views.nbytes), ctx.device_alloc(spr.nbytes), ctx.device_alloc(8), ctx.device_alloc(RING * FB)
ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, 200], np.int32))
def frame(i):
    r = i % RING
    ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, 200, dout + r * FB)
for lanes in (1, 4):
    ctx.set_option(pm3d._lib.OPT_LANES, lanes)
    for i in range(48): frame(i)
    ctx.sync()
    n = 2000
    ctx.timer_start(0); t0 = time.perf_counter()
    for i in range(n): frame(i)
    t1 = time.perf_counter(); ctx.timer_stop(0)
    ms = ctx.timer_ms(0)
    print(f"lanes {lanes}: host enqueue {1e6*(t1-t0)/n:.1f} us/frame, device {1e3*ms/n:.1f} us/frame")
