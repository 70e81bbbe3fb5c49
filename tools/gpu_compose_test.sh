#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_present_gpu.py tests/test_sprite_list_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | tail -5
run() {
  timeout 600 python bench.py --steps $1 --warmup 5 --no-cpu --no-sweeps > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
  python - <<'PY'
import json
l=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print('  value', round(l['value']), 'single', round(l['single_lane']['value']), 'graph', round(l.get('graph_replay',{}).get('value',0)), 'compose_us', round(l['roofline']['kernel_ms']*1e3,1), 'front_us', round(list(l['roofline']['other_kernels_ms'].values())[0]*1e3,1))
PY
}
echo "steps 2000"; run 2000
echo "steps 20"; run 20
OVERLAP=0 python tools/prof_batch.py 4096 5
