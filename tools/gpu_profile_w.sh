mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 1 --no-cpu --samples 4096 --tips 20000 --workload weighted"
$W > gpurun_out/plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted -s 2 -c 1 -o gpurun_out/prof_weighted_v3 $W > gpurun_out/ncu_fw3.log 2>&1
tail -2 gpurun_out/ncu_fw3.log
