#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_present_gpu.py tests/test_sprite_list_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | tail -5
OVERLAP=0 python tools/exp_frame.py 2>&1 | tail -7
timeout 600 python bench.py --steps 400 --warmup 20 --no-cpu --no-sweeps > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, l['single_lane']['value'], l['roofline']['kernel_ms'], l['roofline']['other_kernels_ms'])
PY
