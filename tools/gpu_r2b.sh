# round 2, second GPU pass: warp-aggregated scheduler + ncu source profile of the thread-per-seed kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/exp_r2.py all 2>&1 | tee gpurun_out/r02b_exp.txt | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_t1 -s 2 -c 2 -f -o gpurun_out/r02b_prof_t1_step python tools/profile_t1.py step > gpurun_out/prof_t1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_t1 -s 2 -c 2 -f -o gpurun_out/r02b_prof_t1_speed python tools/profile_t1.py speed >> gpurun_out/prof_t1.log 2>&1
tail -3 gpurun_out/prof_t1.log
ls -la gpurun_out
