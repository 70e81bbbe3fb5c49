#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
python tools/exp_front_stats.py 2>&1 | tail -30
