#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_sprite_list_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head -20
OVERLAP=0 python tools/exp_frame.py 2>&1 | tail -7
for j in 0 1; do
OVERLAP=0 DDA_JUMP=$j python tools/prof_batch.py 4096 5
OVERLAP=0 DDA_JUMP=$j python tools/prof_batch.py 512 5
done
