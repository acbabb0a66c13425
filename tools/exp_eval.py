"""Evaluator roofline launch (4 Mi Panda configurations, all outputs) for the library given as argv[1]."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob
ob.LIB_PATH = sys.argv[1]
dev = torch.device("cuda", 0)
for name, B in (("panda", 1 << 22), ("ur5", 1 << 22), ("snake20", 1 << 20)):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device=dev) for x in r.joint_limits()]
    q = torch.rand((B, n), dtype=torch.float64, device=dev) * (ub - lb) + lb
    tg = r.eval_batch(torch.rand((B, n), dtype=torch.float64, device=dev) * (ub - lb) + lb, want=("ee",))["ee"]
    ts = []
    for i in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = r.eval_batch(q, tg); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
        del out
    ms = float(np.mean(ts)); by = 8 * (8 * n + 17)
    print(f"{name} B={B}: {ms:.4f} ms  {B * by / ms / 1e6:.1f} GB/s  frac {B * by / ms / 1e6 / 6456.8:.3f}", flush=True)
