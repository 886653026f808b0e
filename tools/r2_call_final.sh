# Round-2 evidence on one GPU: tests, the bench lines (ours and the reference arm), launch lists and full ncu
# captures of the dominant kernels at the bench configuration (each after the same command exited 0 without ncu).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 600 gpurun_out/r2_bench_ref.json
timeout 600 python bench.py --workload unweighted --configs none --steps 5 > gpurun_out/r2_bench_unweighted.json 2> gpurun_out/r2_bench_u.err; tail -c 200 gpurun_out/r2_bench_unweighted.json
W="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none"
U="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none --workload unweighted"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_weighted_cfg2.csv $W > gpurun_out/r2_ncu_lw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 3 -c 1 -f -o gpurun_out/r2_prof_isect_weighted $W > gpurun_out/r2_ncu_fw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_mask -s 3 -c 1 -f -o gpurun_out/r2_prof_mask $W > gpurun_out/r2_ncu_fm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_unweighted.csv $U > gpurun_out/r2_ncu_lu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_utips -s 3 -c 1 -f -o gpurun_out/r2_prof_utips $U > gpurun_out/r2_ncu_fu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_unweighted_lut -s 3 -c 1 -f -o gpurun_out/r2_prof_lut $U > gpurun_out/r2_ncu_fl.log 2>&1
for f in gpurun_out/r2_ncu_f*.log gpurun_out/r2_ncu_l*.log; do tail -n 1 $f | cut -c1-160; done
ls -la gpurun_out/r2_prof_*.ncu-rep | tail -8
