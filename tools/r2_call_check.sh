mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_pytest_gpu3.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu3.log
timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu > gpurun_out/r2_bench_check.json 2> gpurun_out/r2_bench_check.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_check.json').read().strip().splitlines()[-1])
print("value",d['value'],"ms",d['ms_per_step'],"embed",d['embed_ms'],"pair",d['pair_ms'],d['prologue_ms']['sparse_rows_and_skip_prep'],"e2e",d['e2e']['seconds'],"api",d['api_e2e']['seconds'] if d.get('api_e2e') else None, d['parity']['max_rel_err_all_ranks'])
P
