# eight (and four) GPUs of one box: the driver's scaling run, product arm, with the peer-to-peer exchange; NCCL for comparison
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --nccl-exchange --headline-only > gpurun_out/r02_bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; tail -3 gpurun_out/bench_n${N}_nccl.err
python - <<PY
import json
for f in ("r02_bench_n$N.json","r02_bench_n${N}_nccl.json"):
    d=json.loads(open('gpurun_out/'+f).read())
    print(f, "value",d['value'],"ms/pass",d['ms_per_pass'],"e2e",d['e2e']['value'],d['e2e']['ms_per_pass'],"exchange",d.get('exchange'))
    if d.get('configs',{}).get('config5'): print(d['configs']['config5'])
    if d.get('per_target'): print({k:v for k,v in d['per_target'].items() if k in('value','all_ranks_sum','ms_per_call')})
PY
