import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, pm3d
from pm3d import workloads as W
from exp_frame import run
ctx = pm3d.Context(0)
ctx.set_option(pm3d._lib.OPT_OVERLAP, int(os.environ.get("OVERLAP", "1")))
c2 = W.config_c2(n_poses=8)
run(ctx, c2, 1920, 1080, "c2 1080p 200 sprites", reps=16)
