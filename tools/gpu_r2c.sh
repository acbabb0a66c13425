set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -25
timeout 600 python tools/exp_r2.py all 2>&1 | tee gpurun_out/r02c_exp.txt | tail -40
