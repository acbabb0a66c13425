"""Device-side timing of eval_kernel (HBM roofline kernel) for several robots."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

def run(name, B, want=("ee", "jac", "f", "grad"), reps=5):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
    q = torch.rand((B, n), dtype=torch.float64, device="cuda") * (ub - lb) + lb
    tg = r.eval_batch(torch.rand((B, n), dtype=torch.float64, device="cuda") * (ub - lb) + lb, want=("ee",))["ee"]
    ts = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = r.eval_batch(q, tg, want=want); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
        del out
    ms = float(np.mean(ts))
    bytes_per = 8 * n + 64 + (64 if "ee" in want else 0) + (48 * n if "jac" in want else 0) + (8 if "f" in want else 0) + (8 * n if "grad" in want else 0)
    print(f"{name} n={n} B={B} want={','.join(want)}: {ms:.3f} ms  {B/ms*1e3:.3e} eval/s  {B*bytes_per/ms/1e6:.1f} GB/s  frac={B*bytes_per/ms/1e6/6541.8:.3f}")

if __name__ == "__main__":
    run("panda", 1 << 22)
    run("ur5", 1 << 22)
    run("snake20", 1 << 20)
    run("panda", 1 << 22, want=("ee",))
    run("panda", 1 << 22, want=("ee", "jac"))
