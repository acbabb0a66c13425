# round 2, first GPU pass: parity of the rewritten thread-per-seed kernel + variant / scheduler experiments
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python tools/exp_r2.py all 2>&1 | tee gpurun_out/r02a_exp.txt | tail -40
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/r02a_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/r02a_bench_n1.json; tail -5 gpurun_out/bench_n1.err
