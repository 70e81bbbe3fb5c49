#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_los_gpu.py tests/test_comm_gpu.py -x -q 2>&1 | tail -4
CMD="python tools/exp_shard.py 64 2097152 10 | grep c5; python tools/exp_shard.py 64 16777216 5 | grep c5; python tools/exp_shard.py 64 262144 10 | grep c5" bash tools/gpu_variants.sh
