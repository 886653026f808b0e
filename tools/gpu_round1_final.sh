# Round-1 evidence run: tests, benches (ours + reference arm), ncu launch lists and full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_weighted_cfg2.json 2> gpurun_out/bench_weighted_cfg2.err; tail -c 600 gpurun_out/bench_weighted_cfg2.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_cfg2.json 2> gpurun_out/bench_reference_cfg2.err; tail -c 400 gpurun_out/bench_reference_cfg2.json
timeout 900 python bench.py --workload unweighted > gpurun_out/bench_unweighted_10k_100k.json 2> gpurun_out/bench_unweighted.err; tail -c 600 gpurun_out/bench_unweighted_10k_100k.json
timeout 900 python bench.py --workload braycurtis > gpurun_out/bench_braycurtis_10k_50k.json 2> gpurun_out/bench_bc.err; tail -c 300 gpurun_out/bench_braycurtis_10k_50k.json
timeout 900 python bench.py --workload jaccard > gpurun_out/bench_jaccard_10k_50k.json 2> gpurun_out/bench_jc.err; tail -c 300 gpurun_out/bench_jaccard_10k_50k.json
W="python bench.py --steps 1 --warmup 1 --no-cpu --samples 4096 --tips 20000 --workload weighted"
U="python bench.py --steps 1 --warmup 1 --no-cpu --samples 4096 --tips 20000 --workload unweighted"
$W > gpurun_out/plain_w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_weighted.csv $W > gpurun_out/ncu_lw.log 2>&1
$U > gpurun_out/plain_u.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_unweighted.csv $U > gpurun_out/ncu_lu.log 2>&1
$W > gpurun_out/plain_w2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted -s 2 -c 1 -o gpurun_out/prof_weighted $W > gpurun_out/ncu_fw.log 2>&1
$U > gpurun_out/plain_u2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_unweighted_lut -s 1 -c 1 -o gpurun_out/prof_unweighted $U > gpurun_out/ncu_fu.log 2>&1
$W > gpurun_out/plain_w3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_propagate_weighted -c 1 -o gpurun_out/prof_propagate $W > gpurun_out/ncu_fp.log 2>&1
tail -n 2 gpurun_out/ncu_fw.log; tail -n 2 gpurun_out/ncu_fu.log; tail -n 2 gpurun_out/ncu_fp.log
