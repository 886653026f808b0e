mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 1 --no-cpu"
$W > gpurun_out/plain_cfg2_mask.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_mask -s 30 -c 1 -o gpurun_out/prof_weighted_mask_cfg2 $W > gpurun_out/ncu_cfg2_mask.log 2>&1
tail -2 gpurun_out/ncu_cfg2_mask.log
