"""Device-side timing of the batched diff_ik kernel."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

for name, B in (("ur3e", 1 << 20), ("panda", 1 << 20)):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
    x0 = torch.rand((B, n), dtype=torch.float64, device="cuda") * (ub - lb) + lb
    V = torch.rand((B, 6), dtype=torch.float64, device="cuda")
    vm = torch.ones((n,), dtype=torch.float64, device="cuda")
    ts = []
    for i in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); al, v, st = r.diff_ik_batch(x0, V, vm); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(f"diff_ik {name} n={n} B={B}: {ms:.3f} ms  {B/ms*1e3:.3e} configurations/s  solved={float((st==1).double().mean()):.5f}  "
          f"alpha mean={float(al.mean()):.3f}  HBM {B*(16*n+56+12)/ms/1e6:.0f} GB/s algorithmic")
