"""Per-warp task timings of one compose launch (library built with -DPM3D_COMPOSE_STATS, e.g. tools/build_variant.sh stats
-DPM3D_COMPOSE_STATS; PM3D_COMPOSE_LOG=path python tools/prof_batch.py 4096 1):
    python tools/exp_tasklog.py gpurun_out/tasklog.bin"""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
t0, t1, sm, row = a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2], a[:, 3]
ok = t1 > 0
t0, t1, sm, row = t0[ok], t1[ok], sm[ok], row[ok]
base = t0.min(); dur = (t1 - t0) / 1e3; end = (t1 - base) / 1e3; start = (t0 - base) / 1e3
print(f"{len(dur)} tasks, launch span {end.max():.1f} us, sum of task time {dur.sum() / 1e3:.1f} ms = {dur.sum() / end.max():.0f} warps busy on average")
print("task duration us: mean %.1f median %.1f p90 %.1f p99 %.1f max %.1f" % (dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
print("last task starts at %.1f us; tasks ending in the last 10%% of the span: %d" % (start.max(), (end > 0.9 * end.max()).sum()))
for lo in range(0, 200, 16):
    m = row == lo
    if m.any(): print(f"  first row {lo:3d}: n {m.sum():6d} mean {dur[m].mean():7.1f} us  max {dur[m].max():7.1f}")
# busy warps over time
edges = np.linspace(0, end.max(), 21)
for i in range(20):
    lo, hi = edges[i], edges[i + 1]
    busy = (np.minimum(end, hi) - np.maximum(start, lo)).clip(0).sum() / (hi - lo)
    print(f"  {lo:7.0f}-{hi:7.0f} us: {busy:7.0f} warps busy")
