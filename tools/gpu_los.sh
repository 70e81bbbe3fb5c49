#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_los_gpu.py tests/test_move_gpu.py tests/test_comm_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/exp_los.py 2>&1 | tail -12
