mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'.')
PY
W="python bench.py --steps 1 --warmup 1 --no-cpu"
$W > gpurun_out/plain_tips.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_pair_tips -s 1 -c 1 -f -o gpurun_out/prof_tips $W > gpurun_out/ncu_tips.log 2>&1
tail -2 gpurun_out/ncu_tips.log | cut -c1-200
