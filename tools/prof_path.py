"""Two FindPath batches (for ncu): 1024 queries on proving_grounds, targets within 60 cells."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
scene, g = W.proving_grounds_scene()
ctx.upload_scene(scene)
q = W.path_queries(g["levels"][0], 1024, 5, max_manhattan=60)
for _ in range(2):
    pts, cnt = ctx.find_path_batch(q, 1024)
print("paths", int((cnt > 0).sum()), "points", int(cnt[cnt > 0].sum()))
