"""c3 (4K) sweep of bench.py alone, for tuning overrides (PM3D_COMPOSE_ROWS / PM3D_COMPOSE_MINB): tools only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import bench
import pm3d
ctx = pm3d.Context(0)
r = bench.sweep_c3(ctx)
for k in ("levels_3", "levels_8"):
    print(os.environ.get("PM3D_COMPOSE_ROWS", "auto"), os.environ.get("PM3D_COMPOSE_MINB", "auto"), k,
          {a: round(b, 4) for a, b in r[k].items() if a in ("ms_per_frame", "compose_ms", "front_ms")})
