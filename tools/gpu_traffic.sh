#!/bin/bash
# steady-state DRAM traffic of the c2 compose launches: two metrics = one pass, no replay, caches left alone, so the
# stores of the 24-frame ring (199 MB > L2) evict each other as they do in the bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_compose -s 60 -c 96 --csv --log-file gpurun_out/r2_traffic_c2.csv python bench.py --steps 200 --warmup 10 --no-cpu --no-sweeps > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log
