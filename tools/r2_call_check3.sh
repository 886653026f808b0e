mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
python tools/r2_lazy_probe.py 2>&1 | tail -2
timeout 600 python tools/fuzz_gpu.py 2>&1 | tail -3
