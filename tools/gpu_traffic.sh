# DRAM traffic of the dominant kernel at the bench configuration (one launch, ncu --set full)
mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 1 --no-cpu"
$W > gpurun_out/plain_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_v3 -s 30 -c 1 -o gpurun_out/prof_weighted_cfg2 $W > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/ncu_cfg2.log
U="python bench.py --steps 1 --warmup 1 --no-cpu --workload unweighted"
$U > gpurun_out/plain_cfg2u.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_unweighted_lut -s 8 -c 1 -o gpurun_out/prof_unweighted_cfg2 $U > gpurun_out/ncu_cfg2u.log 2>&1
tail -2 gpurun_out/ncu_cfg2u.log
