"""Experiment: Speed-mode target batches (BASELINE config 3 shape) -- where does the thread-per-seed kernel stand
relative to its steady-state evaluation rate?"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

def run(name, T, R, mode="speed", chunks=0, tile=1, reps=3, phased=False):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
    g = torch.Generator(device="cuda").manual_seed(42)
    qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    cfg = ob.SolverConfig(solution_mode=mode, max_time=0.0, max_restarts=R)
    best = 1e9
    for i in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        q, f, st, ex = r.ik_batch(cfg, tg, x0, restarts=R, tile=tile, chunks=chunks, stats=True, phased=phased)
        b.record(); torch.cuda.synchronize()
        if i: best = min(best, a.elapsed_time(b))
    cnt = ex["counters"].cpu().numpy()
    ok = float(cfg.is_success(st.cpu().numpy()).mean())
    print(f"{name} T={T} R={R} {mode} phased={phased}: {best:.3f} ms solves/s={T*ok/best*1e3:.3e} ok={ok:.5f} attempts/target={cnt[0]/T:.2f} "
          f"evals/attempt={cnt[1]/cnt[0]:.2f} evals/s={cnt[1]/best*1e3:.3e}")

import os
if os.environ.get("OPTIK_EXP") == "phased":
    for T in (1 << 20, 1 << 18, 1 << 16):
        run("panda", T, 32, phased=True)
    run("ur5", 1 << 20, 32, phased=True)
    sys.exit(0)
for ph in (False, True):
    run("ur5", 1 << 20, 32, phased=ph)
    run("panda", 1 << 20, 32, phased=ph)
    run("panda", 1 << 18, 32, phased=ph)
    run("panda", 1 << 16, 32, phased=ph)
    run("panda", 1 << 14, 32, phased=ph)
run("panda", 1 << 20, 32, mode="quality")
run("panda", 1 << 18, 32, mode="quality")
run("panda", 1 << 16, 256, mode="quality")
run("panda", 1 << 20, 1)
