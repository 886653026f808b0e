# ncu captures for the two dominant kernels (small shapes, one GPU).
mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 1 --no-cpu --samples 4096 --tips 20000 --workload weighted"
U="python bench.py --steps 1 --warmup 1 --no-cpu --samples 4096 --tips 20000 --workload unweighted"
$W > gpurun_out/plain_w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_weighted.csv $W > gpurun_out/ncu_lw.log 2>&1
$U > gpurun_out/plain_u.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_unweighted.csv $U > gpurun_out/ncu_lu.log 2>&1
$W > gpurun_out/plain_w2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted -s 2 -c 2 -o gpurun_out/prof_weighted $W > gpurun_out/ncu_fw.log 2>&1
$U > gpurun_out/plain_u2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_unweighted_lut -s 1 -c 1 -o gpurun_out/prof_unweighted $U > gpurun_out/ncu_fu.log 2>&1
$W > gpurun_out/plain_w3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_propagate_weighted -c 1 -o gpurun_out/prof_propagate $W > gpurun_out/ncu_fp.log 2>&1
tail -2 gpurun_out/plain_w.log gpurun_out/plain_u.log | cut -c1-600
tail -2 gpurun_out/ncu_fw.log gpurun_out/ncu_fu.log gpurun_out/ncu_fp.log
