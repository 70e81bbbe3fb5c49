#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/prof_los.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_los -s 4 -c 2 -o gpurun_out/los_full -f python tools/prof_los.py > gpurun_out/los_full.log 2>&1
tail -3 gpurun_out/los_full.log
