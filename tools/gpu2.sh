# two GPUs: the restart-sharded headline with the peer-to-peer exchange and with the NCCL all-gather
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -x -q -m gpu -k "peer_exchange" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/r02_bench_n2.json; tail -8 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-exchange --headline-only > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; tail -c 600 gpurun_out/r02_bench_n2_nccl.json; tail -3 gpurun_out/bench_n2_nccl.err
