set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
python bench.py --tile 32 --steps 50 --no-cpu-baseline > gpurun_out/bench_n1_tile32.json 2>&1; tail -c 1500 gpurun_out/bench_n1_tile32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 4 -o gpurun_out/prof_solve python tools/profile_target.py > gpurun_out/prof_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -o gpurun_out/prof_eval python tools/profile_target.py > gpurun_out/prof_eval.log 2>&1
ls -la gpurun_out
