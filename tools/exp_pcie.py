"""PCIe ceiling probe (tools only): pinned D2H / H2D bandwidth through pm3d_memcpy_* for frame-sized chunks."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np
import pm3d

ctx = pm3d.Context(0)
for nbytes in (8294400, 4 * 8294400, 64 << 20):
    d = ctx.device_alloc(nbytes)
    h = ctx.host_alloc(nbytes)
    for name, fn in (("d2h", lambda: ctx.d2h(h, d)), ("h2d", lambda: ctx.h2d(d, h))):
        for _ in range(3): fn()
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name} {nbytes/1e6:8.1f} MB pinned, sync per copy: {dt*1e6:8.1f} us = {nbytes/dt/1e9:6.1f} GB/s")
    ctx.device_free(d)
