"""BASELINE config 5 on N GPUs (torchrun): Panda, SolutionMode::Quality, 256 restarts x 1 Mi targets sharded by target,
one NCCL all-gather assembling every target's result on every rank.  Device-side timing, max over ranks.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5_multi.py [T_total] [R]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optik_b200 as ob  # noqa: E402
from optik_b200 import dist as obd  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
T_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
r = ob.Robot.named("panda")
r.set_device(local)
n = r.num_positions()
lo, hi = obd.shard_range(T_total, rank, world)
T = hi - lo
lb, ub = [torch.tensor(x, dtype=torch.float64, device=dev) for x in r.joint_limits()]
g = torch.Generator(device=dev).manual_seed(42 + rank)
qstar = torch.rand((T, n), dtype=torch.float64, device=dev, generator=g) * (ub - lb) + lb
x0 = (torch.rand((T, n), dtype=torch.float64, device=dev, generator=g) * (ub - lb) + lb).contiguous()
targets = r.eval_batch(qstar, want=("ee",))["ee"].contiguous()
cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
times = []
for it in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    q, f, st = obd.ik_batch_target_sharded(r, cfg, targets, x0, R, rank=rank, world=world)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
ms = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
# success gate on the local shard: re-evaluate with the evaluator kernel
ql = q[lo:hi] if q.shape[0] == T_total else q
fl = r.eval_batch(ql.contiguous(), targets, want=("f",))["f"]
stl = st[lo:hi] if st.shape[0] == T_total else st
ok = (stl == 1) & (fl < cfg.tol_f) & ((ql >= lb) & (ql <= ub)).all(dim=1)
cnt = torch.tensor([float(ok.sum())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"config": "5: Panda Quality, %d restarts x %d targets, target-sharded over %d GPU(s), all-gather of results" % (R, T_total, world),
                      "n_gpus": world, "ms": float(ms), "solves_per_s": float(cnt) / float(ms) * 1e3, "success": float(cnt) / T_total,
                      "attempts_per_s": T_total * R / float(ms) * 1e3, "gathered_rows_on_every_rank": int(q.shape[0])}))
if world > 1:
    dist.destroy_process_group()
