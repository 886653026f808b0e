mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1200 python bench.py --workload unweighted --samples 100000 --tips 500000 --lam 0.004 --emulate-shards 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_cfg4_shard0of8.json 2> gpurun_out/bench_cfg4.err; tail -c 1500 gpurun_out/bench_cfg4_shard0of8.json; tail -3 gpurun_out/bench_cfg4.err
timeout 1200 python bench.py --workload weighted_unnorm --tree caterpillar --samples 25000 --tips 1000000 --lam 0.002 --emulate-shards 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_cfg5_shard0of8.json 2> gpurun_out/bench_cfg5.err; tail -c 1500 gpurun_out/bench_cfg5_shard0of8.json; tail -3 gpurun_out/bench_cfg5.err
