mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
SMALL="python bench.py --steps 1 --warmup 1 --no-cpu --samples 4096 --tips 20000"
$SMALL > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
$SMALL > gpurun_out/plain_small2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted -s 2 -c 2 -o gpurun_out/prof_weighted_r1 $SMALL > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
