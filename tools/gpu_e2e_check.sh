for ke in 50 200 200 50; do
OPTIK_BENCH_E2E_STEPS=$ke python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('Ke', d['e2e']['steps'], 'value %.4g e2e %.4g blocking %.4g ms/step %.4f'%(d['value'], d['e2e']['value'], d['e2e']['blocking_call_value'], d['e2e']['ms_per_step']))"
done
