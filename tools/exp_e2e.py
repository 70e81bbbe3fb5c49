import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
c2 = W.config_c2(n_poses=8)
ctx.upload_scene(c2["scene"]); ctx.set_view(1920, 1080, 1, 200)
views, spr, off = c2["views"], c2["sprites"], c2["offsets"]
FB = 1920*1080*4
host = [ctx.host_alloc(FB, np.uint8, (1,1080,1920,4)) for _ in range(3)]
dbuf = ctx.device_alloc(FB)
def T(f, n=200):
    f(); ctx.sync(); t=time.perf_counter()
    for i in range(n): f(i)
    ctx.sync(); return (time.perf_counter()-t)/n*1e6
print("d2h 8.3MB pinned, sync each: %.1f us" % T(lambda i=0: ctx.d2h(host[0], dbuf)))
print("render no frame copy (sync): %.1f us" % T(lambda i=0: ctx.render(views[i%8:i%8+1], spr[off[i%8]:off[i%8+1]], want_frame=False)))
print("render + frame copy (sync): %.1f us" % T(lambda i=0: ctx.render(views[i%8:i%8+1], spr[off[i%8]:off[i%8+1]], out=host[i%3])))
print("render no frame, no_sync:   %.1f us" % T(lambda i=0: ctx.render(views[i%8:i%8+1], spr[off[i%8]:off[i%8+1]], want_frame=False, no_sync=True)))
ctx.set_option(pm3d._lib.OPT_PIPELINE, 3)
print("render + frame, pipelined:  %.1f us" % T(lambda i=0: ctx.render(views[i%8:i%8+1], spr[off[i%8]:off[i%8+1]], out=host[i%3], no_sync=True)))
# host cost of the call alone: time only the enqueue loop (no sync inside), frames pile up on the device
def H(f, n=200):
    f(); ctx.sync(); t = time.perf_counter()
    for i in range(n): f(i)
    dt = time.perf_counter() - t; ctx.sync(); return dt / n * 1e6
print("host enqueue cost, pipelined: %.1f us/call" % H(lambda i=0: ctx.render(views[i%8:i%8+1], spr[off[i%8]:off[i%8+1]], out=host[i%3], no_sync=True)))
for depth in (2, 4):
    ctx.set_option(pm3d._lib.OPT_PIPELINE, depth)
    hh = [ctx.host_alloc(FB, np.uint8, (1,1080,1920,4)) for _ in range(depth)]
    print("pipelined depth %d: %.1f us" % (depth, T(lambda i=0: ctx.render(views[i%8:i%8+1], spr[off[i%8]:off[i%8+1]], out=hh[i%depth], no_sync=True))))
