mkdir -p gpurun_out
python tools/r2_cfg4_shard_run.py > gpurun_out/r2_cfg4_shard_plain.log 2>&1 || { tail -3 gpurun_out/r2_cfg4_shard_plain.log; exit 1; }
tail -1 gpurun_out/r2_cfg4_shard_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 1 -c 1 -f -o gpurun_out/r2_prof_isect_cfg4 python tools/r2_cfg4_shard_run.py > gpurun_out/r2_ncu_cfg4.log 2>&1; tail -1 gpurun_out/r2_ncu_cfg4.log
