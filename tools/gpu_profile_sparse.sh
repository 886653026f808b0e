mkdir -p gpurun_out
S="python bench.py --steps 1 --warmup 1 --no-cpu --samples 20000 --tips 50000 --workload braycurtis"
$S > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_sparse -s 1 -c 1 -o gpurun_out/prof_sparse $S > gpurun_out/ncu_s.log 2>&1
tail -2 gpurun_out/ncu_s.log
