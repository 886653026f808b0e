mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; tail -c 300 gpurun_out/r2_bench_8gpu.err
timeout 600 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --configs cfg3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; tail -c 200 gpurun_out/r2_bench_4gpu.err
timeout 600 $TR --nproc-per-node 2 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --configs cfg3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -c 200 gpurun_out/r2_bench_2gpu.err
python - <<'P'
import json
for f in ("r2_bench_8gpu","r2_bench_4gpu","r2_bench_2gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["prologue_ms"], d["e2e"]["value"], d["e2e"]["seconds"], d["parity"]["max_rel_err_all_ranks"])
        for k,v in (d.get("configs") or {}).items():
            print("  ", k, v["value"], v["ms_per_step"], v["prologue_ms"].get("exchange"), v["e2e"]["value"], v["parity"]["max_rel_err_all_ranks"], v["roofline"].get("frac"))
    except Exception as e: print(f, "ERR", e)
P
