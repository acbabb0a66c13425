set -x
python tools/probe_speed.py 2>&1 | head -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_n1_b.json 2>gpurun_out/bench_n1_b.err; tail -c 600 gpurun_out/bench_n1_b.json | head -c 600; tail -3 gpurun_out/bench_n1_b.err
