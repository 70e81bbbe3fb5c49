#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for k in 1 2; do
timeout 600 python bench.py --steps 2000 --warmup 5 --no-cpu --no-sweeps > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/g_bench.json").read().strip().splitlines()[-1])
print("RESULT value", round(l["value"]), "single", round(l["single_lane"]["value"]), "compose_us", round(l["roofline"]["kernel_ms"]*1e3,1), "front_us", round(list(l["roofline"]["other_kernels_ms"].values())[0]*1e3,1))
PY
done
