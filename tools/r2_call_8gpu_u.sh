mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 --configs none --no-api --workload unweighted > gpurun_out/r2_bench_8gpu_unweighted.json 2> gpurun_out/r2_bench_8gpu_unweighted.err; tail -c 300 gpurun_out/r2_bench_8gpu_unweighted.err
python - <<'P'
import json
for f in ("r2_bench_8gpu_unweighted",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["prologue_ms"], d["e2e"]["value"], d["parity"]["max_rel_err_all_ranks"])
    except Exception as e: print(f, "ERR", e)
P
