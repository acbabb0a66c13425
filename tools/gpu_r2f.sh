set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 6000 gpurun_out/r02f_bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02f_bench_ref_n1.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/r02f_bench_ref_n1.json
