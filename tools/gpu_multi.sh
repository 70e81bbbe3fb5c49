#!/bin/bash
# multi-GPU check: comm / group tests + the bench under torchrun on all visible GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_comm_gpu.py tests/test_sprite_list_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"
grep -a "pm3d-sweeps" gpurun_out/bench_${N}gpu.err | tail -2
tail -c 600 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','scaling')})
print(l.get('e2e',{}).get('value'))
PY
