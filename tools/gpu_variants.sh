#!/bin/bash
# run CMD (default: the c4 batch timings) once per library variant under variants/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
cp pixelmek-3d_b200/libpm3d.so /tmp/libpm3d_default.so
for v in variants/libpm3d_*.so; do
  cp $v pixelmek-3d_b200/libpm3d.so
  echo "== $v"
  if [ -n "$CMD" ]; then bash -c "$CMD"; else
  OVERLAP=0 python tools/prof_batch.py 4096 5; OVERLAP=0 python tools/prof_batch.py 512 5; fi
done
cp /tmp/libpm3d_default.so pixelmek-3d_b200/libpm3d.so
