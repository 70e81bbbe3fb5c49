"""Batched-views run for ncu (throughput mode of the compose kernel): N views of 320x200, resident."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np
import pm3d
from pm3d import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = W.config_c4(n_views=n)
ctx = pm3d.Context(0)
ctx.set_option(pm3d._lib.OPT_OVERLAP, int(os.environ.get('OVERLAP', '1')))
ctx.set_option(pm3d._lib.OPT_DDA_JUMP, int(os.environ.get('DDA_JUMP', '1')))
ctx.upload_scene(cfg["scene"]); ctx.set_view(320, 200, n, len(cfg["sprites"]))
v, s, o = cfg["views"], cfg["sprites"], cfg["offsets"]
dv, ds, do = ctx.device_alloc(v.nbytes), ctx.device_alloc(s.nbytes), ctx.device_alloc(o.nbytes)
dout = ctx.device_alloc(n * 320 * 200 * 4)
ctx.h2d(dv, v); ctx.h2d(ds, s); ctx.h2d(do, o)
for _ in range(2): ctx.render_resident(dv, n, ds, do, len(s), dout)
ctx.sync()
ctx.ktimer_begin(reps); ctx.timer_start(0)
for _ in range(reps): ctx.render_resident(dv, n, ds, do, len(s), dout)
ctx.timer_stop(0); ms = ctx.timer_ms(0) / reps
kt = ctx.ktimer_collect(reps)
px = n * 320 * 200
print(f"{n} views: {ms*1e3:.1f} us/sweep, front {kt[:,0].mean()*1e3:.1f} us, compose {kt[:,2].mean()*1e3:.1f} us = {kt[:,2].mean()*1e9/px:.2f} ps/px, {px*4/kt[:,2].mean()/1e6:.0f} GB/s")
