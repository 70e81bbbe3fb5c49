import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, pm3d
from pm3d import workloads as W
from exp_frame import run
ctx = pm3d.Context(0)
c1 = W.config_c1(); c1["offsets"] = np.array([0, len(c1["sprites"])], np.int32)
run(ctx, c1, 640, 360, "c1 640x360 408 sprites", reps=10)
