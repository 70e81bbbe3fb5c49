import os, sys, ctypes as C
os.environ["PM3D_WARP_CLOCKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, pm3d
from pm3d import workloads as W
ctx = pm3d.Context(0)
c2 = W.config_c2(n_poses=8)
views, spr, off = c2["views"], c2["sprites"], c2["offsets"]
ctx.upload_scene(c2["scene"]); ctx.set_view(1920, 1080, 1, 200)
ctx.set_option(pm3d._lib.OPT_OVERLAP, 0)
for pose in (0, 3):
    for rep in range(3):
        ctx.render(views[pose:pose+1], spr[off[pose]:off[pose+1]], want_frame=False)
    ctx.lib.pm3d_dbg_clocks.restype = C.c_void_p
    ptr = ctx.lib.pm3d_dbg_clocks()
    n = 30 * 135 * 4
    buf = np.zeros((n, 4), dtype=np.int64)
    ctx.d2h(buf, ptr)
    dur, t0, meta, sm = buf[:, 0], buf[:, 1], buf[:, 2], buf[:, 3]
    nl = meta >> 32; row = meta & 0xffffffff
    ok = dur > 0
    print(f"pose {pose}: warps {ok.sum()}  warp dur mean {dur[ok].mean():.0f} p50 {np.percentile(dur[ok],50):.0f} p90 {np.percentile(dur[ok],90):.0f} p99 {np.percentile(dur[ok],99):.0f} max {dur[ok].max()}")
    # by row band
    for lo in range(0, 1080, 120):
        m = ok & (row >= lo) & (row < lo + 120)
        if m.sum(): print(f"   rows {lo:4d}-{lo+119:4d}: mean {dur[m].mean():7.0f} max {dur[m].max():7d}  n_list mean {nl[m].mean():5.1f} max {nl[m].max():3d}")
    top = np.argsort(-dur)[:8]
    M = (1 << 20) - 1
    print("   slowest (row, n_list, total | prologue, f2b, bg, repaint clk):")
    for i in top: print("     ", int(row[i]), int(nl[i]), int(dur[i]), "|", int(t0[i] & M), int((t0[i] >> 20) & M), int((t0[i] >> 40) & M), int((sm[i] >> 44) & M))
    med = np.argsort(dur)[len(dur)//2]
    print("   median warp:", int(row[med]), int(nl[med]), int(dur[med]), "|", int(t0[med] & M), int((t0[med] >> 20) & M), int((t0[med] >> 40) & M), int((sm[med] >> 44) & M))
