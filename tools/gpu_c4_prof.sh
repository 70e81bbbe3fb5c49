#!/bin/bash
# c4 compose / front under ncu --set full with source (throughput mode of the compose kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
OVERLAP=0 python tools/prof_batch.py 4096 3 || exit 1
OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 2 -c 2 -o gpurun_out/prof_c4_s4 -f python tools/prof_batch.py 4096 2 > gpurun_out/ncu_c4_s4.log 2>&1
tail -2 gpurun_out/ncu_c4_s4.log
ls -la gpurun_out/*.ncu-rep
