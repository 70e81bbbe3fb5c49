import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np
import pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
c = W.config_c2(w=640, h=360, n_poses=6)
n = len(c["views"])
offsets = (np.arange(n + 1) * len(c["sprites"])).astype(np.int32)
spr = np.concatenate([c["sprites"]] * n)
ctx.upload_scene(c["scene"])
ctx.set_option(pm3d._lib.OPT_KEEP_COLUMNS, 1)
ctx.set_view(640, 360, n, len(spr))
res = {}
for j in (0, 1):
    ctx.set_option(pm3d._lib.OPT_DDA_JUMP, j)
    ctx.render(c["views"], spr, offsets)
    res[j] = [ctx.download_columns(i)[0] for i in range(n)]
for v in range(n):
    a, b = res[0][v], res[1][v]
    bad = np.nonzero((a["map_x"] != b["map_x"]) | (a["map_y"] != b["map_y"]) | (a["side"] != b["side"]) | (a["hit"] != b["hit"]))
    print("view", v, "pose", c["poses"][v], "bad", len(bad[0]), "of", a.size)
    for lv, x in list(zip(*bad))[:6]:
        print("  level", lv, "x", x, "seq", a[lv, x]["map_x"], a[lv, x]["map_y"], a[lv, x]["side"], a[lv, x]["hit"], "jump", b[lv, x]["map_x"], b[lv, x]["map_y"], b[lv, x]["side"], b[lv, x]["hit"])
