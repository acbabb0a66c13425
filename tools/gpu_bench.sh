set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 4000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --tile 32 --steps 50 --no-cpu-baseline > gpurun_out/bench_n1_tile32.json 2>gpurun_out/bench_n1_tile32.err; tail -c 1800 gpurun_out/bench_n1_tile32.json; tail -3 gpurun_out/bench_n1_tile32.err
python -c "import __graft_entry__ as g; g.smoke()"
