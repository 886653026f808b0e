mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python tools/r2_api_probe.py 2>&1 | tail -7
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print("value",d['value'],"ms",d['ms_per_step'],"e2e",d['e2e']['value'],d['e2e']['seconds'],"api",d['api_e2e']['seconds'])
P
