# launch list + full captures of the two pair kernels at BASELINE config 2 (default bench command)
mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 1 --no-cpu"
$W > gpurun_out/plain_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_weighted_cfg2.csv $W > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_tips -s 1 -c 1 -f -o gpurun_out/prof_tips $W > gpurun_out/ncu_tips.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_mask -s 1 -c 1 -f -o gpurun_out/prof_mask $W > gpurun_out/ncu_mask.log 2>&1
tail -1 gpurun_out/ncu_mask.log | cut -c1-120
