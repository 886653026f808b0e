# Round-2 ncu evidence: launch lists of one bench step and full captures of the dominant kernels at the
# bench configuration (each only after the same command has exited 0 without ncu).
mkdir -p gpurun_out
W="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none"
U="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none --workload unweighted"
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --configs none --workload braycurtis --samples 50000 --tips 50000"
$W > gpurun_out/r2_plain_w.log 2>&1 || { tail -5 gpurun_out/r2_plain_w.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_weighted_cfg2.csv $W > gpurun_out/r2_ncu_lw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 3 -c 1 -f -o gpurun_out/r2_prof_isect_weighted $W > gpurun_out/r2_ncu_fw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_weighted_mask -s 3 -c 1 -f -o gpurun_out/r2_prof_mask $W > gpurun_out/r2_ncu_fm.log 2>&1
$U > gpurun_out/r2_plain_u.log 2>&1 || { tail -5 gpurun_out/r2_plain_u.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_unweighted.csv $U > gpurun_out/r2_ncu_lu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 3 -c 1 -f -o gpurun_out/r2_prof_isect_unweighted $U > gpurun_out/r2_ncu_fu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_unweighted_lut -s 3 -c 1 -f -o gpurun_out/r2_prof_lut $U > gpurun_out/r2_ncu_fl.log 2>&1
$B > gpurun_out/r2_plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 3 -c 1 -f -o gpurun_out/r2_prof_isect_features $B > gpurun_out/r2_ncu_fb.log 2>&1
for f in gpurun_out/r2_ncu_f*.log; do tail -n 1 $f | cut -c1-160; done
ls -la gpurun_out/*.ncu-rep | tail -8
