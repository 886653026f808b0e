mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --configs none --no-api > gpurun_out/r2_x2_on.json 2> gpurun_out/r2_x2_on.err; tail -c 300 gpurun_out/r2_x2_on.err
python - <<'P'
import json
for f in ("on",):
    try:
        d=json.loads(open(f"gpurun_out/r2_x2_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["prologue_ms"], d["e2e"], d["parity"]["max_rel_err_all_ranks"])
    except Exception as e: print(f, "ERR", e)
P
