import sys; sys.path.insert(0,"tools"); sys.path.insert(0,".")
import optik_b200 as ob
ob.LIB_PATH=sys.argv[1]
import exp_r2
for T in (5000, 8192, 16384, 30000):
    for R in (2, 32):
        exp_r2.batch("panda", T, R)
exp_r2.batch("panda", 1<<16, 32)
exp_r2.batch("ur3e", 1<<14, 100)
# dynamic chains == static schedule (same per-target answer), for the library under test
import numpy as np, torch
for name, T, R in (("panda", 20000, 16), ("ur5", 50000, 32), ("panda", 6000, 24)):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
    g = torch.Generator(device="cuda").manual_seed(5)
    qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q1, f1, s1, e1 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True, static=True, chunks=1)
    same = True
    for rep in range(3):
        q2, f2, s2, e2 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True)
        ok = torch.as_tensor(cfg.is_success(s1.cpu().numpy()), device="cuda")
        same &= bool(torch.equal(s1, s2) and torch.equal(q1[ok], q2[ok]) and torch.equal(e1["restart"][ok], e2["restart"][ok]) and torch.equal(q1[~ok], q2[~ok]))
    print(f"{name} T={T} R={R}: dynamic == static: {same}", flush=True)
