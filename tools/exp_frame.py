"""Ad-hoc timing experiments for the frame path (not part of the product or the tests)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np
import pm3d
from pm3d import workloads as W

def run(ctx, cfg, w, h, label, sprites=True, reps=200, ring=8):
    views, spr, off = cfg["views"], cfg["sprites"], cfg["offsets"]
    ns = int(off[1] - off[0]) if sprites else 0
    ctx.upload_scene(cfg["scene"]); ctx.set_view(w, h, 1, max(ns, 1))
    n = len(views)
    dv = ctx.device_alloc(views.nbytes); ds = ctx.device_alloc(max(spr.nbytes, 64)); do = ctx.device_alloc(8)
    dout = ctx.device_alloc(ring * w * h * 4)
    ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, ns], np.int32))
    def frame(i):
        r = i % n
        ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, ns, dout + (i % ring) * w * h * 4)
    for i in range(20): frame(i)
    ctx.sync()
    ctx.ktimer_begin(reps); ctx.timer_start(0)
    for i in range(reps): frame(i)
    ctx.timer_stop(0); ms = ctx.timer_ms(0) / reps
    kt = ctx.ktimer_collect(reps)
    print(f"{label:40s} frame {ms*1e3:7.1f} us | cast {kt[:,0].mean()*1e3:6.1f} setup {kt[:,1].mean()*1e3:6.1f} compose {kt[:,2].mean()*1e3:6.1f} (min {kt[:,2].min()*1e3:5.1f} max {kt[:,2].max()*1e3:5.1f})")
    for p in (dv, ds, do, dout): ctx.device_free(p)

if __name__ == "__main__":
    ctx = pm3d.Context(0)
    ctx.set_option(pm3d._lib.OPT_DDA_JUMP, int(os.environ.get("DDA_JUMP", "1")))
    ctx.set_option(pm3d._lib.OPT_OVERLAP, int(os.environ.get("OVERLAP", "1")))
    print("DDA_JUMP =", os.environ.get("DDA_JUMP", "1"), "OVERLAP =", os.environ.get("OVERLAP", "1"))
    c2 = W.config_c2(n_poses=8)
    run(ctx, c2, 1920, 1080, "c2 1080p 200 sprites")
    run(ctx, c2, 1920, 1080, "c2 1080p no sprites", sprites=False)
    c2s = W.config_c2(w=640, h=360, n_poses=8)
    run(ctx, c2s, 640, 360, "c2 640x360 200 sprites")
    c3 = W.config_c3(levels=3); c3["offsets"] = np.array([0, 0], np.int32)
    run(ctx, c3, 3840, 2160, "c3 4K 3 levels", sprites=False, reps=50, ring=2)
    c3 = W.config_c3(levels=8); c3["offsets"] = np.array([0, 0], np.int32)
    run(ctx, c3, 3840, 2160, "c3 4K 8 levels", sprites=False, reps=50, ring=2)
    c1 = W.config_c1(); c1["offsets"] = np.array([0, len(c1["sprites"])], np.int32)
    run(ctx, c1, 640, 360, "c1 640x360 408 sprites")
