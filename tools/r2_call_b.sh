mkdir -p gpurun_out
timeout 900 python tools/r2_isect_probe.py > gpurun_out/r2_isect_probe.log 2>&1; cat gpurun_out/r2_isect_probe.log | tail -60
W="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none"
$W > gpurun_out/r2_plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 3 -c 1 -f -o gpurun_out/r2_prof_isect_weighted $W > gpurun_out/r2_ncu_fw.log 2>&1
tail -2 gpurun_out/r2_ncu_fw.log | cut -c1-200
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
