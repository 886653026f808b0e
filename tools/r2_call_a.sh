mkdir -p gpurun_out
python tools/r2_microbench.py > gpurun_out/r2_microbench.log 2>&1; cat gpurun_out/r2_microbench.log
timeout 900 python tools/r2_isect_probe.py > gpurun_out/r2_isect_probe.log 2>&1; cat gpurun_out/r2_isect_probe.log | tail -40
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; tail -40 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
