mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "intersection" > gpurun_out/r2_pytest_flat.log 2>&1; tail -3 gpurun_out/r2_pytest_flat.log
timeout 900 python tools/r2_isect_probe.py flat > gpurun_out/r2_isect_probe_flat.log 2>&1; cat gpurun_out/r2_isect_probe_flat.log
