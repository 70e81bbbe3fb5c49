#!/bin/bash
# GPU call A: the parity suite + a short bench of the current build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
grep -E "^E|passed|failed|rc=" gpurun_out/a_pytest.log | tail -15
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/a_bench.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/a_bench.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','single_lane','e2e','roofline')})
    for k,v in l.get('extra',{}).items(): print(k, v)
except Exception as e: print('parse failed', e)
PY
