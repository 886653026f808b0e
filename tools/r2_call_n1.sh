mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -3 gpurun_out/r2_bench_final.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print("value",d['value'],"ms",d['ms_per_step'],"e2e",d['e2e']['value'],d['e2e']['seconds'],"api",d['api_e2e']['seconds'],d['api_e2e']['split'])
s=d['roofline'].get('sparse_row_kernel') or d['roofline']; print(s.get('shared_memory_pipe'))
P
