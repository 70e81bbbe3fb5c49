"""Timing of the presentation kernel (tools only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
for (w, h, ow, oh, sc) in ((960, 540, 1920, 1080, 0.5), (1920, 1080, 1920, 1080, 1.0), (1920, 1080, 3840, 2160, 0.5)):
    cfg = W.config_c2(w, h, 60, n_poses=1)
    ctx.upload_scene(cfg["scene"]); ctx.set_view(w, h, 1, 60)
    ctx.render(cfg["views"], cfg["sprites"], cfg["offsets"], want_frame=False)
    ctx.set_present(ow, oh, sc)
    ring = 8; fb = ow * oh * 4
    d_out = ctx.device_alloc(ring * fb)
    for name, mode in (("copy", 0), ("crt", 1), ("curve", 2)):
        for i in range(5): ctx.present(mode, 1, d_out=d_out + (i % ring) * fb)
        ctx.sync(); reps = 100
        ctx.timer_start(5)
        for i in range(reps): ctx.present(mode, 1, d_out=d_out + (i % ring) * fb)
        ctx.timer_stop(5); ms = ctx.timer_ms(5) / reps
        print(f"{w}x{h} -> {ow}x{oh} {name:6s} {ms*1e3:7.1f} us  {fb/ms/1e6:7.0f} GB/s out")
    ctx.device_free(d_out)
