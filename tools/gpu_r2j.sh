# re-entry check of HEAD: GPU tests, smoke, bench N=1 (product arm)
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/bench_n1.err; head -c 600 gpurun_out/r02j_bench_n1.json; tail -3 gpurun_out/bench_n1.err
