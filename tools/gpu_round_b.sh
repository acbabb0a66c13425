# round-1 (second session) artefacts: tests, smoke, both bench arms, launch list, ncu --set full of the step's kernels and
# of the evaluator, per-config table, probes
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01b_bench_ref_n1.json 2> gpurun_out/bench_ref.err; head -c 400 gpurun_out/r01b_bench_ref_n1.json
timeout 600 python bench.py > gpurun_out/r01b_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 4500 gpurun_out/r01b_bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"solve_t1|select_kernel" -s 2 -c 3 -f -o gpurun_out/r01b_prof_step python tools/profile_target.py solve > gpurun_out/prof_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -f -o gpurun_out/r01b_prof_eval python tools/profile_target.py eval > gpurun_out/prof_eval.log 2>&1
tail -2 gpurun_out/prof_step.log gpurun_out/prof_eval.log
timeout 600 python tools/configs_table.py > gpurun_out/r01b_configs_table.md 2> gpurun_out/configs.err; cat gpurun_out/r01b_configs_table.md; tail -3 gpurun_out/configs.err
timeout 300 python tools/probe_eval.py 2>&1 | tail -6
timeout 300 python tools/probe_diffik.py 2>&1 | tail -3
ls -la gpurun_out
