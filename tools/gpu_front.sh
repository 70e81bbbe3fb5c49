#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
OVERLAP=0 python tools/exp_frame.py 2>&1 | tail -7
for n in 512 4096; do OVERLAP=0 python tools/prof_batch.py $n 5; done
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_sprite_list_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head -20
