"""Stress of the dynamic Speed chains: many repetitions of batches whose targets are shared between lanes (record word,
restart counters, tickets, in-warp speculation), each compared with the static schedule's per-target answer."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0
for name, T, R in (("panda", 9500, 32), ("panda", 12000, 6), ("panda", 37000, 32), ("panda", 38000, 32), ("panda", 90000, 16),
                   ("ur5", 20000, 32), ("ur5", 300000, 32), ("ur3e", 10000, 64), ("ur3e", 60000, 100),
                   # edges of the schedule: the dynamic threshold, exactly the resident lanes, one and two restarts
                   ("panda", 9472, 1), ("panda", 9473, 2), ("panda", 37887, 3), ("panda", 37888, 1), ("panda", 37889, 5),
                   ("panda", 50000, 1), ("ur5", 12629, 2), ("ur5", 56832, 3)):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
    g = torch.Generator(device="cuda").manual_seed(T)
    qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q1, f1, s1, e1 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True, static=True, chunks=1)
    ok = torch.as_tensor(cfg.is_success(s1.cpu().numpy()), device="cuda")
    fails = 0
    for rep in range(REPS):
        q2, f2, s2, e2 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True, variant=1 + rep % 2)
        same = bool(torch.equal(s1, s2) and torch.equal(q1[ok], q2[ok]) and torch.equal(f1[ok], f2[ok]) and
                    torch.equal(e1["restart"][ok], e2["restart"][ok]) and torch.equal(q1[~ok], q2[~ok]))
        fails += (not same)
    bad += fails
    print(f"{name} T={T} R={R}: {REPS - fails}/{REPS} repetitions identical to the static schedule (solved {float(ok.double().mean()):.5f})", flush=True)
print("STRESS", "OK" if bad == 0 else f"FAILED ({bad})")
