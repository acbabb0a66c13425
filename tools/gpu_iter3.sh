timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.4f e2e %.4g blocking %.4g e2e ms/step %.4f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_call_value'], d['e2e']['ms_per_step']))"
python tools/exp_speed_batch.py 2>&1 | grep -E "T=1048576" | cut -c1-150
