# launch list + one full capture of the masked weighted kernel at BASELINE config 2
mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 1 --no-cpu"
$W > gpurun_out/plain_cfg2_mask.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_weighted_cfg2_skip.csv $W > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_mask -s 15 -c 1 -o gpurun_out/prof_weighted_mask_cfg2 -f $W > gpurun_out/ncu_cfg2_mask.log 2>&1
tail -2 gpurun_out/ncu_cfg2_mask.log | cut -c1-200
