# usage: bash tools/gpu_scale_w.sh N   (default weighted workload, weak scaling, torchrun; ours only)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_weighted_${N}gpu.json 2> gpurun_out/bench_weighted_${N}gpu.err
tail -2 gpurun_out/bench_weighted_${N}gpu.err
python - <<PY
import json
for l in open("gpurun_out/bench_weighted_${N}gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["config"]["samples"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"])
PY
