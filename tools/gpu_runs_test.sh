#!/bin/bash
# parity of the >= 8-row compose shapes (runs, persistent grid) + the c4 / c2 timings of every library variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "runs or every_compose or c4 or overflow or ragged or c1 or c2" 2>&1 | tail -5
CMD="OVERLAP=0 python tools/prof_batch.py 4096 5; python bench.py --steps 400 --warmup 5 --no-cpu --no-sweeps 2>/dev/null | python tools/bench_brief.py" bash tools/gpu_variants.sh
