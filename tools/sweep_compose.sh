#!/bin/bash
# tuning sweep of the compose kernel's rows-per-warp / min-blocks variants (tools only)
for cfg in "2 6" "4 8" "8 6" "8 8"; do set -- $cfg
  echo "== rows=$1 minb=$2"
  PM3D_COMPOSE_ROWS=$1 PM3D_COMPOSE_MINB=$2 python tools/exp_frame.py | tail -6
  PM3D_COMPOSE_ROWS=$1 PM3D_COMPOSE_MINB=$2 python tools/prof_batch.py 4096 3
done
