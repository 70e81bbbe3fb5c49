# The ncu captures behind profiles/ (run through gpurun on one B200; every command first runs once without ncu).
set -x
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/plain_bench_r1c.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench_r1c.csv python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/ncu_launches_c.log 2>&1
python tools/prof_clutter.py 4096 3 && ncu --set full --clock-control none --import-source on -k regex:k_clutter -c 1 -o gpurun_out/prof_clutter_r1 -f python tools/prof_clutter.py 4096 1 > gpurun_out/ncu_clutter.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 20 -c 4 -o gpurun_out/prof_bench_c2_r1b -f python bench.py --steps 40 --warmup 5 --no-cpu --no-sweeps > gpurun_out/ncu_c2b.log 2>&1
python tools/exp_c3.py && ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 8 -c 2 -o gpurun_out/prof_c3_r1 -f python tools/exp_c3.py > gpurun_out/ncu_c3.log 2>&1
python tools/prof_batch.py 512 2 && ncu --set full --clock-control none --import-source on -k regex:"k_compose|k_front" -s 2 -c 2 -o gpurun_out/prof_c4_r1g -f python tools/prof_batch.py 512 2 > gpurun_out/ncu_c4g.log 2>&1
python tools/exp_los.py && ncu --set full --clock-control none --import-source on -k regex:k_los -s 2 -c 1 -o gpurun_out/prof_los_r1g -f python tools/exp_los.py > gpurun_out/ncu_losg.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
