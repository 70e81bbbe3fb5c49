#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sprite_list_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed|Error" | tail -8
for g in 0 1; do
PM3D_BENCH_GRAPH=$g timeout 600 python bench.py --steps 2000 --warmup 50 --no-cpu --no-sweeps > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "graph=$g bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, 'single', l['single_lane']['value'])
PY
done
tail -3 gpurun_out/g_bench.err
