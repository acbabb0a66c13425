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
x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
for _ in range(3):
    r.ik_attempts(cfg, targets[0], x0, R, tile=1, best=True)
torch.cuda.synchronize()
print("done")
