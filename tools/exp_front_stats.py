"""Front-kernel statistics (jump fallbacks, slowest cast warp / view CTA in cycles) per pose; needs a library built
with EXTRA=-DPM3D_FRONT_STATS (tools/gpu_front_stats.sh)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
ctx.set_option(pm3d._lib.OPT_OVERLAP, 0)
lib = C.CDLL(os.path.join(ROOT, "pixelmek-3d_b200", "libpm3d.so"))
lib.pm3d_debug_front_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
def stats(reset=1):
    out = (C.c_ulonglong * 12)()
    lib.pm3d_debug_front_stats(ctx.h, out, reset)
    return list(out)
def run(cfg, w, h, label, sprites=True):
    views, spr, off = cfg["views"], cfg["sprites"], cfg["offsets"]
    ns = int(off[1] - off[0]) if sprites else 0
    ctx.upload_scene(cfg["scene"]); ctx.set_view(w, h, 1, max(ns, 1))
    n = len(views)
    dv = ctx.device_alloc(views.nbytes); ds = ctx.device_alloc(max(spr.nbytes, 64)); do = ctx.device_alloc(8)
    dout = ctx.device_alloc(w * h * 4)
    ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, ns], np.int32))
    for r in range(n):
        for _ in range(3):
            ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, ns, dout)
        ctx.sync(); stats(1)
        ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, ns, dout)
        ctx.sync(); st = stats(1)
        print(f"{label} pose {r}: jump rays {st[1]}, fallbacks {st[0]}, slowest cast CTA-thread {st[2]} cycles, view CTA {st[3]} cycles; jump iterations max {st[4]} mean {st[7]/max(st[1],1):.1f}, long-jump tries {st[5]/max(st[1],1):.1f} per ray; slowest loop {st[6]} cycles; per iteration: jump {st[9]/max(st[7],1):.0f} probe {st[10]/max(st[7],1):.0f} cycles")
    for p in (dv, ds, do, dout): ctx.device_free(p)
c2 = W.config_c2(n_poses=8)
run(c2, 1920, 1080, "c2")
run(c2, 1920, 1080, "c2 nosprites", sprites=False)
c3 = W.config_c3(levels=3); c3["offsets"] = np.array([0, 0], np.int32)
run(c3, 3840, 2160, "c3", sprites=False)
