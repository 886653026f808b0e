mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --configs none --no-api --samples 10240 > gpurun_out/r2_x2_even.json 2> gpurun_out/r2_x2_even.err; tail -c 600 gpurun_out/r2_x2_even.err
timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --configs none --no-api --samples 10240 --workload unweighted > gpurun_out/r2_x2_even_u.json 2> gpurun_out/r2_x2_even_u.err; tail -c 300 gpurun_out/r2_x2_even_u.err
python - <<'P'
import json
for f in ("r2_x2_even","r2_x2_even_u"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["prologue_ms"], d["e2e"]["value"], d["parity"]["max_rel_err_all_ranks"])
    except Exception as e: print(f, "ERR", e)
P
