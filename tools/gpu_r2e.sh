set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "twin or dynamic or batch or config" 2>&1 | tail -4
timeout 600 python tools/exp_r2.py batch 2>&1 | tee gpurun_out/r02e_exp.txt | tail -40
timeout 600 python tools/exp_r2.py step 2>&1 | tee -a gpurun_out/r02e_exp.txt | tail -8
