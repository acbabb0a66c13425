"""Experiment: thread-per-seed kernel, one Panda target -- throughput vs number of seeds and vs grid size
(how much of the 65 536-seed step is tail, how much is steady state)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

r = ob.Robot.named("panda")
lb, ub = map(np.array, r.joint_limits())
rng = np.random.default_rng(42)
dev = torch.device("cuda", 0)
qstar = torch.from_numpy(rng.uniform(lb, ub, size=(4, 7))).to(dev)
targets = r.eval_batch(qstar, want=("ee",))["ee"].contiguous()
x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev).reshape(1, 7)

def run(R, blocks=0, chunks=0, reps=5):
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        q, f, st, ex = r.ik_batch(cfg, targets[:1], x0, restarts=R, tile=1, chunks=chunks or R, blocks=blocks, stats=True)
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    cnt = ex["counters"].cpu().numpy()
    ms = float(np.median(ts))
    print(f"R={R:8d} blocks={blocks:4d}: {ms:8.3f} ms  attempts/s={cnt[0]/ms*1e3:.3e} evals/s={cnt[1]/ms*1e3:.3e} conv/s={cnt[2]/ms*1e3:.3e} evals/att={cnt[1]/cnt[0]:.2f}")

for R in (18944, 37888, 65536, 131072, 262144, 1048576):
    run(R)
for b in (74, 148, 222, 296):
    run(65536, blocks=b)
for b in (148, 296):
    run(1048576, blocks=b)
