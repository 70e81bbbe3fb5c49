"""c2 resident frame rate on N GPUs at once, one plain process per GPU (tools only)."""
import os, sys, time, multiprocessing as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))

def worker(dev, barrier, q):
    import numpy as np, pm3d
    from pm3d import workloads as W
    ctx = pm3d.Context(dev)
    c2 = W.config_c2(n_poses=24, seed=12345 + 1000 * dev)
    views, spr, off = c2["views"], c2["sprites"], c2["offsets"]
    ctx.upload_scene(c2["scene"]); ctx.set_view(1920, 1080, 1, 200)
    FB = 1920 * 1080 * 4; ring = 24
    dv = ctx.device_alloc(views.nbytes); ds = ctx.device_alloc(spr.nbytes); do = ctx.device_alloc(8); dout = ctx.device_alloc(ring * FB)
    ctx.h2d(dv, views); ctx.h2d(ds, spr); ctx.h2d(do, np.array([0, 200], np.int32))
    def frame(i):
        r = i % ring
        ctx.render_resident(dv + r * 104, 1, ds + int(off[r]) * 64, do, 200, dout + r * FB)
    for i in range(50): frame(i)
    ctx.sync(); barrier.wait()
    reps = 2000
    t0 = time.perf_counter()
    ctx.timer_start(0)
    for i in range(reps): frame(i)
    t_host = time.perf_counter() - t0
    ctx.timer_stop(0); ms = ctx.timer_ms(0) / reps
    q.put((dev, 1e3 / ms, t_host / reps * 1e6))

if __name__ == "__main__":
    import pm3d
    n = pm3d.load().pm3d_device_count()
    mp.set_start_method("spawn", force=True)
    for group in ([0], list(range(n))):
        b = mp.Barrier(len(group)); q = mp.Queue()
        ps = [mp.Process(target=worker, args=(d, b, q)) for d in group]
        [p.start() for p in ps]; [p.join() for p in ps]
        res = sorted(q.get() for _ in group)
        print(f"{len(group)} GPU(s):", ", ".join(f"gpu{d} {g:.0f} fps (host {h:.1f} us/frame)" for d, g, h in res), f"| total {sum(g for _, g, _ in res):.0f}")
