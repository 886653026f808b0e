mkdir -p gpurun_out
nvidia-smi -L | head -8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $T bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_weighted_8gpu.json 2> gpurun_out/bench_weighted_8gpu.err
tail -c 1800 gpurun_out/bench_weighted_8gpu.json; tail -3 gpurun_out/bench_weighted_8gpu.err
timeout 1500 $T bench.py --gpus 8 --workload unweighted --samples 100000 --tips 500000 --lam 0.004 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_cfg4_8gpu.json 2> gpurun_out/bench_cfg4_8gpu.err
tail -c 1800 gpurun_out/bench_cfg4_8gpu.json; tail -3 gpurun_out/bench_cfg4_8gpu.err
