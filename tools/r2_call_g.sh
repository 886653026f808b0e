mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -c 600 gpurun_out/r2_bench_g.json; tail -4 gpurun_out/r2_bench_g.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 900 gpurun_out/r2_bench_ref.json
timeout 600 python bench.py --workload unweighted --configs none --steps 5 > gpurun_out/r2_bench_unweighted.json 2> gpurun_out/r2_bench_u.err; tail -c 300 gpurun_out/r2_bench_unweighted.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
