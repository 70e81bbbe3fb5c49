#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -x -q 2>&1 | grep -E "^E|passed|failed|Error" | head -20
timeout 600 python tools/exp_path.py 2>&1 | tail -5
