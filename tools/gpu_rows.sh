#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --steps $1 --warmup 5 --no-cpu --no-sweeps > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
  python - <<'PY'
import json
l=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print('  value', round(l['value']), 'single', round(l['single_lane']['value']), 'graph', round(l.get('graph_replay',{}).get('value',0)), 'compose_us', round(l['roofline']['kernel_ms']*1e3,1))
PY
}
echo "default 2000"; run 2000
echo "default 20"; run 20
echo "rows4 minb6 2000"; PM3D_COMPOSE_ROWS=4 PM3D_COMPOSE_MINB=6 run 2000
echo "rows4 minb8 2000"; PM3D_COMPOSE_ROWS=4 PM3D_COMPOSE_MINB=8 run 2000
echo "rows4 minb8 20"; PM3D_COMPOSE_ROWS=4 PM3D_COMPOSE_MINB=8 run 20
echo "rows8 2000"; PM3D_COMPOSE_ROWS=8 run 2000
