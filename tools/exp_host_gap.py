"""How much of a mid-size Speed batch's event time is host launch latency?  One call between two events vs ten calls
queued back to back (the GPU never waits for the host after the first)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob
r = ob.Robot.named("panda")
n = 7
lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
g = torch.Generator(device="cuda").manual_seed(42)
for T in (4096, 16384, 65536, 262144):
    qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    for R in (2, 32):
        cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
        out = None
        for _ in range(3):
            out = r.ik_batch(cfg, tg, x0, restarts=R)
        torch.cuda.synchronize()
        one = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r.ik_batch(cfg, tg, x0, restarts=R); b.record(); torch.cuda.synchronize()
            one.append(a.elapsed_time(b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(10):
            r.ik_batch(cfg, tg, x0, restarts=R)
        b.record()
        t_host = (time.perf_counter() - t0) / 10
        torch.cuda.synchronize()
        print(f"T={T} R={R}: single call {min(one):.3f} ms   back-to-back {a.elapsed_time(b) / 10:.3f} ms/call   host enqueue {t_host * 1e3:.3f} ms/call", flush=True)
