mkdir -p gpurun_out
for v in "1 4" "1 -8"; do
  set -- $v
  tag=g$2
  python tools/r2_variant_run.py weighted_unifrac_normalized $1 $2 > gpurun_out/r2_var_$tag.log 2>&1 || { tail -3 gpurun_out/r2_var_$tag.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:k_pair_isect -s 2 -c 1 -f -o gpurun_out/r2_prof_isect_w_$tag python tools/r2_variant_run.py weighted_unifrac_normalized $1 $2 > gpurun_out/r2_ncu_var_$tag.log 2>&1
  tail -2 gpurun_out/r2_ncu_var_$tag.log | cut -c1-200
done
ls -la gpurun_out/r2_prof_isect_w_*
