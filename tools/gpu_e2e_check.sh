for d in 2 3 4 6; do
OPTIK_BENCH_E2E_DEPTH=$d python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('depth', d['e2e']['pipeline_depth'], 'value %.4g e2e %.4g blocking %.4g ms/step %.4f'%(d['value'], d['e2e']['value'], d['e2e']['blocking_call_value'], d['e2e']['ms_per_step']))"
done
