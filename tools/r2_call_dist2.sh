mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "distributed_build or devices_argument or partial_upload" > gpurun_out/r2_pytest_dist.log 2>&1; tail -15 gpurun_out/r2_pytest_dist.log
