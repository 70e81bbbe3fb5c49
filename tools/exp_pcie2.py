"""Aggregate PCIe D2H probe (tools only): one process per GPU copying 8.3 MB frames concurrently."""
import os, sys, time, multiprocessing as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))

def worker(dev, n_dev, barrier, q):
    import numpy as np, pm3d
    ctx = pm3d.Context(dev)
    nbytes = 8294400
    d = ctx.device_alloc(nbytes); h = ctx.host_alloc(nbytes)
    for _ in range(5): ctx.d2h(h, d)
    barrier.wait()
    t0 = time.perf_counter(); reps = 300
    for _ in range(reps): ctx.d2h(h, d)
    dt = time.perf_counter() - t0
    q.put((dev, nbytes * reps / dt / 1e9))

if __name__ == "__main__":
    import pm3d
    n = pm3d.load().pm3d_device_count()
    for group in ([0], list(range(n))):
        mp.set_start_method("spawn", force=True)
        b = mp.Barrier(len(group)); q = mp.Queue()
        ps = [mp.Process(target=worker, args=(d, len(group), b, q)) for d in group]
        [p.start() for p in ps]; [p.join() for p in ps]
        res = sorted(q.get() for _ in group)
        print(f"{len(group)} GPU(s) copying at once:", ", ".join(f"gpu{d} {g:.1f} GB/s" for d, g in res), f"| total {sum(g for _, g in res):.1f} GB/s")
