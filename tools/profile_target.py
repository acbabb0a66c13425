"""Small launch sequence for ncu captures: the bench step's kernels (solve_t1 / solve<8> / solve<32> + select) on the
bench workload, and eval_kernel at the roofline batch (4 Mi Panda configurations)."""
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
cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=65536)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "solve"):
    for tile in (1, 1, 8, 32):
        r.ik_attempts(cfg, targets[0], x0, 65536, tile=tile, best=True)
    torch.cuda.synchronize()
if which in ("all", "eval"):
    B = 1 << 22
    lb_t, ub_t = torch.from_numpy(lb).to(dev), torch.from_numpy(ub).to(dev)
    q = torch.rand((B, 7), dtype=torch.float64, device=dev) * (ub_t - lb_t) + lb_t
    tg = r.eval_batch(torch.rand((B, 7), dtype=torch.float64, device=dev) * (ub_t - lb_t) + lb_t, want=("ee",))["ee"]
    for _ in range(2):
        r.eval_batch(q, tg)
    torch.cuda.synchronize()
print("done")
