"""Batched FindPath timing: 4096 queries on proving_grounds (neighbourhood targets and cross-map targets) and on the c5 map."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
from oracle import oracle as O
ctx = pm3d.Context(0)
cases = [("proving_grounds <= 60 cells", W.load_golden("proving_grounds")["levels"][0], 60),
         ("proving_grounds any", W.load_golden("proving_grounds")["levels"][0], None),
         ("c5 256x256 10% any", W.synthetic_levels(256, 256, 1, 0.10, 42)[0], None)]
n, cap = 4096, 2048
for label, lvl, mm in cases:
    ctx.upload_scene(W.scene_from_grid(np.ascontiguousarray(lvl)[None], 1))
    q = W.path_queries(lvl, n, 5, max_manhattan=mm)
    d_q, d_p, d_c = ctx.device_alloc(q.nbytes), ctx.device_alloc(n * cap * 16), ctx.device_alloc(n * 4)
    ctx.h2d(d_q, q)
    call = lambda: ctx._check(ctx.lib.pm3d_find_path_batch(ctx.h, C.c_void_p(d_q), n, cap, C.c_void_p(d_p), C.c_void_p(d_c), 7), "path")
    call(); ctx.sync()
    ctx.timer_start(0); call(); ctx.timer_stop(0)
    ms = ctx.timer_ms(0)
    cnt = np.zeros(n, np.int32); ctx.d2h(cnt, d_c)
    t = time.time(); _, wn, cells, exp = O.find_path_batch(lvl, q[:256], cap=cap, threads=os.cpu_count()); cpu = (time.time() - t) / 256
    assert np.array_equal(cnt[:256], wn)
    print(f"{label}: {ms:.1f} ms for {n} queries = {n/ms:.1f} k queries/s; CPU oracle {1/cpu:.0f} queries/s on {os.cpu_count()} threads; mean path {cells.mean():.0f} cells, {exp.mean():.0f} nodes popped", flush=True)
    for p in (d_q, d_p, d_c): ctx.device_free(p)
