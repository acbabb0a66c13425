# round-1 re-entry: validate HEAD on a B200 (tests, smoke, both bench arms) and take a source-level capture of solve_t1
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 4500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_t1 -s 1 -c 1 -o gpurun_out/r01b_prof_t1 python tools/profile_target.py solve > gpurun_out/prof_t1.log 2>&1
tail -2 gpurun_out/prof_t1.log
ls -la gpurun_out
