#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
OVERLAP=0 python tools/prof_c2.py || exit 1
OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:k_compose -s 20 -c 8 -o gpurun_out/compose_c2 -f python tools/prof_c2.py > gpurun_out/compose_c2.log 2>&1
tail -2 gpurun_out/compose_c2.log
