mkdir -p gpurun_out
timeout 900 python tools/r2_isect_probe.py w u f > gpurun_out/r2_isect_probe_e.log 2>&1; cat gpurun_out/r2_isect_probe_e.log | cut -c1-190
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "pipelined or flagged or sparse_rows or feature or fuzz or randomised" > gpurun_out/r2_pytest_e.log 2>&1; tail -3 gpurun_out/r2_pytest_e.log
timeout 900 python bench.py --workload unweighted --samples 100000 --tips 500000 --lam 0.004 --emulate-shards 8 --configs none --no-cpu --no-api --steps 2 --warmup 1 > gpurun_out/r2_cfg4_shard.json 2> gpurun_out/r2_cfg4_shard.err; tail -c 1500 gpurun_out/r2_cfg4_shard.json; tail -3 gpurun_out/r2_cfg4_shard.err
