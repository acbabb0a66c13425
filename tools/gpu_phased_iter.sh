set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "phased or ik_batch or config3 or config5 or edge" 2>&1 | tail -12
timeout 300 python tools/exp_speed_batch.py 2>&1 | tail -14
