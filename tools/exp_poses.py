"""Per-pose kernel times of a c2 ring (tools only): python tools/exp_poses.py SEED"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 12345
ctx = pm3d.Context(0)
c2 = W.config_c2(n_poses=24, seed=seed)
views, spr, off = c2["views"], c2["sprites"], c2["offsets"]
ctx.upload_scene(c2["scene"]); ctx.set_view(1920, 1080, 1, 200)
ctx.set_option(pm3d._lib.OPT_OVERLAP, 0)
FB = 1920 * 1080 * 4
dv = ctx.device_alloc(views.nbytes); ds = ctx.device_alloc(spr.nbytes); do = ctx.device_alloc(8); dout = ctx.device_alloc(2 * FB)
ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, 200], np.int32))
for r in range(24):
    f = lambda: ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, 200, dout)
    for _ in range(3): f()
    ctx.sync(); ctx.ktimer_begin(10)
    for _ in range(10): f()
    kt = ctx.ktimer_collect(10)
    v = views[r]
    print(f"pose {r:2d} pos ({v['pos_x']:6.1f},{v['pos_y']:6.1f}) front {kt[:,0].mean()*1e3:6.1f} us compose {kt[:,2].mean()*1e3:6.1f} us")
