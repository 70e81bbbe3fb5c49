# The ncu captures behind profiles/r2b_* (end of round 2; one B200 through gpurun; every command first runs once without ncu).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
set -x
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/plain_bench_r2b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/ncu_launches_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 20 -c 4 -o gpurun_out/prof_c2_r2b -f python bench.py --steps 40 --warmup 5 --no-cpu --no-sweeps > gpurun_out/ncu_c2_r2b.log 2>&1
OVERLAP=0 python tools/prof_batch.py 4096 2 && OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 2 -c 2 -o gpurun_out/prof_c4_r2b -f python tools/prof_batch.py 4096 2 > gpurun_out/ncu_c4_r2b.log 2>&1
python tools/prof_los.py && ncu --set full --clock-control none --import-source on -k regex:k_los -s 4 -c 2 -o gpurun_out/prof_los_r2b -f python tools/prof_los.py > gpurun_out/ncu_los_r2b.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_compose -s 60 -c 96 --csv --log-file gpurun_out/r2b_traffic_c2.csv python bench.py --steps 200 --warmup 10 --no-cpu --no-sweeps > gpurun_out/ncu_traffic_r2b.log 2>&1
ls -la gpurun_out/*_r2b*
