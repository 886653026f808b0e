mkdir -p gpurun_out
timeout 900 python bench.py --workload braycurtis --samples 50000 --tips 50000 --steps 3 --warmup 3 > gpurun_out/bench_cfg3_braycurtis_50kx50k.json 2> gpurun_out/bench_cfg3_bc.err; tail -c 2200 gpurun_out/bench_cfg3_braycurtis_50kx50k.json; tail -3 gpurun_out/bench_cfg3_bc.err
timeout 900 python bench.py --workload jaccard --samples 50000 --tips 50000 --steps 3 --warmup 3 > gpurun_out/bench_cfg3_jaccard_50kx50k.json 2> gpurun_out/bench_cfg3_j.err; tail -c 1200 gpurun_out/bench_cfg3_jaccard_50kx50k.json; tail -3 gpurun_out/bench_cfg3_j.err
