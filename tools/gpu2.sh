# two GPUs: the restart-sharded headline with the peer-to-peer exchange and with the NCCL all-gather
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --headline-only > gpurun_out/r02_bench_n2.json 2> gpurun_out/bench_n2.err; tail -4 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-exchange --headline-only > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; tail -3 gpurun_out/bench_n2_nccl.err
python - <<'PY'
import json
for f in ("r02_bench_n2.json","r02_bench_n2_nccl.json"):
    d=json.loads(open('gpurun_out/'+f).read())
    print(f, "value",d['value'],"ms/pass",d['ms_per_pass'],"e2e",d['e2e']['value'],d['e2e']['ms_per_pass'],"exchange",d.get('exchange'))
PY
