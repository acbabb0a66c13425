set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -f -o gpurun_out/r01b_prof_eval python tools/profile_target.py eval > gpurun_out/prof_eval.log 2>&1
tail -2 gpurun_out/prof_eval.log
