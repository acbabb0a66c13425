# quick loop for evaluator work: its parity tests + device timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "evaluator or golden or jacobian or edge or chain" 2>&1 | tail -8
timeout 300 python tools/probe_eval.py 2>&1 | tail -8
