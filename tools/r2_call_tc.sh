mkdir -p gpurun_out
timeout 600 python tools/r2_tc_probe.py small > gpurun_out/r2_tc_probe.log 2>&1; tail -8 gpurun_out/r2_tc_probe.log | cut -c1-260
timeout 600 python tools/r2_tc_probe.py cfg2 >> gpurun_out/r2_tc_probe.log 2>&1; tail -8 gpurun_out/r2_tc_probe.log | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "tensor_cores" > gpurun_out/r2_pytest_tc.log 2>&1; tail -12 gpurun_out/r2_pytest_tc.log
