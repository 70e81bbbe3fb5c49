# The ncu captures behind profiles/ (round 2; run through gpurun on one B200; every command first runs once without ncu).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
set -x
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/plain_bench_r2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 20 -c 6 -o gpurun_out/prof_c2_r2 -f python bench.py --steps 40 --warmup 5 --no-cpu --no-sweeps > gpurun_out/ncu_c2_r2.log 2>&1
OVERLAP=0 python tools/prof_batch.py 4096 2 && OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 2 -c 2 -o gpurun_out/prof_c4_r2 -f python tools/prof_batch.py 4096 2 > gpurun_out/ncu_c4_r2.log 2>&1
python tools/prof_los.py && ncu --set full --clock-control none --import-source on -k regex:k_los -s 4 -c 2 -o gpurun_out/prof_los_r2 -f python tools/prof_los.py > gpurun_out/ncu_los_r2.log 2>&1
python tools/prof_path.py && ncu --set full --clock-control none --import-source on -k regex:k_find_path -s 1 -c 1 -o gpurun_out/prof_path_r2 -f python tools/prof_path.py > gpurun_out/ncu_path_r2.log 2>&1
ls -la gpurun_out/*_r2.ncu-rep
