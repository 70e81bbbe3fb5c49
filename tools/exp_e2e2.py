"""e2e pipelined loop on N GPUs at once, one plain process per GPU, no torch (tools only)."""
import os, sys, time, multiprocessing as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))

def worker(dev, barrier, q, import_torch):
    if import_torch:
        import torch
        torch.cuda.set_device(dev)
        torch.zeros(1, device="cuda")
    import numpy as np, pm3d
    from pm3d import workloads as W
    ctx = pm3d.Context(dev)
    c2 = W.config_c2(n_poses=8, seed=12345 + 1000 * dev)
    ctx.upload_scene(c2["scene"]); ctx.set_view(1920, 1080, 1, 200)
    views, spr, off = c2["views"], c2["sprites"], c2["offsets"]
    FB = 1920 * 1080 * 4
    host = [ctx.host_alloc(FB, np.uint8, (1, 1080, 1920, 4)) for _ in range(3)]
    ctx.set_option(pm3d._lib.OPT_PIPELINE, 3)
    f = lambda i: ctx.render(views[i % 8:i % 8 + 1], spr[off[i % 8]:off[i % 8 + 1]], out=host[i % 3], no_sync=True)
    for i in range(5): f(i)
    ctx.sync(); barrier.wait()
    n = 300; t = time.perf_counter()
    for i in range(n): f(i)
    ctx.sync(); dt = time.perf_counter() - t
    q.put((dev, n / dt))

if __name__ == "__main__":
    import pm3d
    n = pm3d.load().pm3d_device_count()
    mp.set_start_method("spawn", force=True)
    for import_torch in (False, True):
        for group in ([0], list(range(n))):
            b = mp.Barrier(len(group)); q = mp.Queue()
            ps = [mp.Process(target=worker, args=(d, b, q, import_torch)) for d in group]
            [p.start() for p in ps]; [p.join() for p in ps]
            res = sorted(q.get() for _ in group)
            print(f"torch={import_torch} {len(group)} GPU(s):", ", ".join(f"gpu{d} {g:.0f} fps" for d, g in res), f"| total {sum(g for _, g in res):.0f}")
