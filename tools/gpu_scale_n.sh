# usage: bash tools/gpu_scale_n.sh N   (default weighted workload, weak scaling, torchrun)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_weighted_${N}gpu.json 2> gpurun_out/bench_weighted_${N}gpu.err
tail -c 900 gpurun_out/bench_weighted_${N}gpu.json; tail -2 gpurun_out/bench_weighted_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 1 --warmup 1 --impl reference > gpurun_out/bench_reference_${N}gpu.json 2>/dev/null; tail -c 300 gpurun_out/bench_reference_${N}gpu.json
