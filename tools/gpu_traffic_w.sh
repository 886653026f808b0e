mkdir -p gpurun_out
nvidia-smi --query-gpu=serial,clocks.sm,clocks.mem,clocks.max.mem,temperature.gpu,temperature.memory,ecc.mode.current --format=csv
W="python bench.py --steps 1 --warmup 1 --no-cpu"
$W > gpurun_out/plain_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_v3 -s 30 -c 1 -o gpurun_out/prof_weighted_cfg2_b $W > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/ncu_cfg2.log
