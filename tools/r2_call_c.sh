mkdir -p gpurun_out
timeout 900 python tools/r2_isect_probe.py w u f > gpurun_out/r2_isect_probe_d.log 2>&1; cat gpurun_out/r2_isect_probe_d.log | tail -40
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "pipelined or flagged or sparse_rows or feature" > gpurun_out/r2_pytest_d.log 2>&1; tail -3 gpurun_out/r2_pytest_d.log
W="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none"
$W > gpurun_out/r2_plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 3 -c 1 -f -o gpurun_out/r2_prof_isect_weighted_d $W > gpurun_out/r2_ncu_fw.log 2>&1
tail -2 gpurun_out/r2_ncu_fw.log | cut -c1-200
