"""ncu target: thread-per-seed kernel on Panda, 256 Ki targets -- (1) Speed R=32 single launch, (2) Speed R=1, (3) Quality R=8."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob
r = ob.Robot.named("panda")
n = 7
T = 1 << 18
lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
g = torch.Generator(device="cuda").manual_seed(42)
qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
for mode, R in (("speed", 32), ("speed", 1), ("quality", 8)):
    cfg = ob.SolverConfig(solution_mode=mode, max_time=0.0, max_restarts=R)
    r.ik_batch(cfg, tg, x0, restarts=R, tile=1, chunks=1)
torch.cuda.synchronize()
print("done")
