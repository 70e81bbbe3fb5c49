"""c2 resident throughput with K contexts (K independent streams) on one GPU, frames dealt round-robin (tools only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np
import pm3d
from pm3d import workloads as W

w, h, ring, ns = 1920, 1080, 24, 200
cfg = W.config_c2(w, h, ns, n_poses=ring)
views, spr, off = cfg["views"], cfg["sprites"], cfg["offsets"]
for K in (1,):
    ctxs = [pm3d.Context(0) for _ in range(K)]
    bufs = []
    for c in ctxs:
        c.upload_scene(cfg["scene"]); c.set_view(w, h, 1, ns)
        dv, ds, do = c.device_alloc(views.nbytes), c.device_alloc(spr.nbytes), c.device_alloc(8)
        dout = c.device_alloc(ring * w * h * 4)
        c.h2d(dv, views); c.h2d(ds, spr); c.h2d(do, np.array([0, ns], np.int32))
        bufs.append((dv, ds, do, dout))
    def frame(i):
        c = ctxs[i % K]; dv, ds, do, dout = bufs[i % K]; r = i % ring
        c.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, ns, dout + r * w * h * 4)
    for i in range(48): frame(i)
    for c in ctxs: c.sync()
    n = 2400
    t0 = time.perf_counter()
    for c in ctxs: c.timer_start(0)
    for i in range(n): frame(i)
    for c in ctxs: c.timer_stop(0)
    ms = max(c.timer_ms(0) for c in ctxs)
    for c in ctxs: c.sync()
    wall = time.perf_counter() - t0
    print(f"K={K}: {n / (ms * 1e-3):8.0f} frames/s by events (max over lanes), {n / wall:8.0f} by wall clock")
    for c in ctxs: c.close()

# the same through ONE ctx with PM3D_OPT_LANES
for lanes in (1, 2, 3, 4):
    c = pm3d.Context(0)
    c.upload_scene(cfg["scene"]); c.set_view(w, h, 1, ns)
    c.set_option(pm3d._lib.OPT_LANES, lanes)
    dv, ds, do = c.device_alloc(views.nbytes), c.device_alloc(spr.nbytes), c.device_alloc(8)
    dout = c.device_alloc(ring * w * h * 4)
    c.h2d(dv, views); c.h2d(ds, spr); c.h2d(do, np.array([0, ns], np.int32))
    def frame1(i):
        r = i % ring
        c.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, ns, dout + r * w * h * 4)
    for i in range(48): frame1(i)
    c.sync()
    n = 2400
    t0 = time.perf_counter()
    c.timer_start(0)
    for i in range(n): frame1(i)
    t_enq = time.perf_counter() - t0
    c.timer_stop(0)
    ms = c.timer_ms(0); c.sync()
    print(f"   host enqueue loop {t_enq / n * 1e6:.1f} us per frame, device {ms / n * 1e3:.1f} us per frame")
    wall = time.perf_counter() - t0
    print(f"lanes={lanes}: {n / (ms * 1e-3):8.0f} frames/s by events, {n / wall:8.0f} by wall clock")
    c.close()
