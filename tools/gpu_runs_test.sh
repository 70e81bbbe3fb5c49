#!/bin/bash
# batch + single-frame timings of every library variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="for n in 1024 512 256; do OVERLAP=0 python tools/prof_batch.py \$n 5; done; python bench.py --steps 400 --warmup 5 --no-cpu --no-sweeps 2>/dev/null | python tools/bench_brief.py" bash tools/gpu_variants.sh
