#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_sprite_list_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | tail -5
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"
grep -a "pm3d-sweeps" gpurun_out/s_bench.err | tail -1
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, l['single_lane']['value'], l['roofline']['kernel_ms'], l['roofline']['other_kernels_ms'])
print(l['extra'].get('astar'))
PY
