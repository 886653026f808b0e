mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --configs cfg3 --no-api > gpurun_out/r2_bench_2gpu_last.json 2> gpurun_out/r2_bench_2gpu_last.err; tail -c 200 gpurun_out/r2_bench_2gpu_last.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_2gpu_last.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["prologue_ms"]["exchange"], d["e2e"]["value"], d["parity"]["max_rel_err_all_ranks"])
for k,v in d["configs"].items(): print("  ", k, v["value"], v["ms_per_step"], v["parity"]["max_rel_err_all_ranks"])
P
