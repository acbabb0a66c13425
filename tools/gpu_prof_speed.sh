set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_t1 -c 3 -f -o gpurun_out/r01b_prof_speed python tools/profile_speed.py > gpurun_out/prof_speed.log 2>&1
tail -2 gpurun_out/prof_speed.log
