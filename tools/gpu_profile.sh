set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_ -s 1 -c 3 -o gpurun_out/r01_prof_solve python tools/profile_target.py solve > gpurun_out/prof_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -o gpurun_out/r01_prof_eval python tools/profile_target.py eval > gpurun_out/prof_eval.log 2>&1
tail -2 gpurun_out/prof_solve.log gpurun_out/prof_eval.log
