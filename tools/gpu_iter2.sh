set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('value %.4g ms/step %.4f e2e %.4g (blocking %.4g) roof %.3f evals/s %.4g succ %.3f' % (d['value'], d['ms_per_step'], d["e2e"]["value"], d["e2e"]["blocking_call_value"], d["roofline"]["frac"], d['roofline_solve']['evals_per_s'], d['success_rate_per_attempt']), d['verified_equals_claimed'], d['oracle_spot_check_ok'])
PY
timeout 300 python tools/exp_speed_batch.py 2>&1 | tail -14 | cut -c1-170
