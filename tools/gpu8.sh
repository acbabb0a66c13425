set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r01b_bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r01b_bench_n$N.json')); print($N, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['blocking_call_value'], d['clocks'])"
tail -2 gpurun_out/bench_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 2>/dev/null > gpurun_out/r01b_bench_ref_n8.json; head -c 300 gpurun_out/r01b_bench_ref_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/config5_multi.py 2>/dev/null | tail -1 | tee gpurun_out/r01b_config5_n8.json
