#!/bin/bash
# what the driver runs at round end: smoke, the bench with its short step count, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
t0=$(date +%s.%N)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$? wall since $t0"
t0=$(date +%s.%N)
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/d_ref.json 2> gpurun_out/d_ref.err; echo "reference rc=$? wall since $t0"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/d_bench.json").read().strip().splitlines()[-1]); r=json.loads(open("gpurun_out/d_ref.json").read().strip().splitlines()[-1])
print("ours", round(l["value"]), "e2e", round(l["e2e"]["value"]), "ref", round(r["value"],1), "ratio e2e", round(l["e2e"]["value"]/r["value"],1), "ratio value", round(l["value"]/r["value"]), "len", len(json.dumps(l)))
print("same config", l["config"]==r["config"], l.get("graph_replay"), l["clocks"])
PY
