set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --headline-only --no-cpu-baseline > gpurun_out/r02g_bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_n1.json').read())
print("value",d['value'],"ms/pass",d['ms_per_pass'],"e2e",d['e2e']['value'],d['e2e']['ms_per_pass'],"blocking",d['e2e']['blocking_call_value'],d['e2e']['blocking_call_ms_median'], d['roofline_solve']['frac'])
PY
tail -3 gpurun_out/bench_n1.err
