# round-2 artefacts: tests, smoke, both bench arms, ncu launch list of the bench command, ncu --set full captures of the
# solve kernels (bench step, Speed batch, Quality batch, snake, single ik) -> gpurun_out/r02_* (summaries go to profiles/)
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/bench_ref.err; head -c 300 gpurun_out/r02_bench_ref_n1.json
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench_n1.err; head -c 400 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --passes 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for w in step speed quality snake single; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solve_t1|solve_kernel" -s 2 -c 2 -f -o gpurun_out/r02_prof_$w python tools/profile_t1.py $w > gpurun_out/prof_$w.log 2>&1; tail -1 gpurun_out/prof_$w.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_eval python tools/profile_target.py eval > gpurun_out/prof_eval.log 2>&1; tail -1 gpurun_out/prof_eval.log
ls -la gpurun_out
