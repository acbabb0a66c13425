"""Launch sequence for ncu captures of the thread-per-seed kernel:
  step     the bench step (65 536 seeds, one Panda target), both variants
  speed    a Speed batch (1 Mi Panda targets, 32 restarts, dynamic chains)
  quality  a Quality batch (262 144 Panda targets x 32 restarts, static jobs)
  snake    BASELINE config 4 (20-DOF snake, 262 144 seeds): thread-per-seed kernel x3, then the tile kernel
  single   four Robot::ik calls (tile kernel with the fused selection, mapped memory)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

r = ob.Robot.named("panda")
lb, ub = map(np.array, r.joint_limits())
rng = np.random.default_rng(42)
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "step"
if which == "step":
    qstar = torch.from_numpy(rng.uniform(lb, ub, size=(4, 7))).to(dev)
    targets = r.eval_batch(qstar, want=("ee",))["ee"].contiguous()
    x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev)
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=65536)
    import json
    for v in (0, 0, 0, 0):  # the library's own choice of layout for this shape (two column rows in shared memory)
        cnt = torch.zeros(3, dtype=torch.int64, device=dev)
        r.ik_attempts(cfg, targets[0], x0, 65536, best=True, variant=v, counters=cnt)
    json.dump({"attempts": int(cnt[0]), "evaluations": int(cnt[1]), "converged": int(cnt[2])}, open("gpurun_out/r02_step_counters.json", "w"))
elif which == "steady":  # 1 Mi seeds to one target: steady state of both layouts (launches 3 and 4 are the ones to capture)
    qstar = torch.from_numpy(rng.uniform(lb, ub, size=(4, 7))).to(dev)
    targets = r.eval_batch(qstar, want=("ee",))["ee"].contiguous()
    x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev)
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=1 << 20)
    for v in (1, 2, 1, 2):
        r.ik_attempts(cfg, targets[0], x0, 1 << 20, best=True, variant=v)
elif which == "mid":  # a mid-size Speed batch (fewer targets than resident lanes: shared chains from the start)
    T = 16384
    lb_t, ub_t = torch.from_numpy(lb).to(dev), torch.from_numpy(ub).to(dev)
    g = torch.Generator(device="cuda").manual_seed(42)
    qs = torch.rand((T, 7), dtype=torch.float64, device=dev, generator=g) * (ub_t - lb_t) + lb_t
    x0s = (torch.rand((T, 7), dtype=torch.float64, device=dev, generator=g) * (ub_t - lb_t) + lb_t).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    for R in (2, 32, 2, 32):
        r.ik_batch(ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R), tg, x0s, restarts=R)
elif which == "snake":
    snake = ob.Robot.named("snake20")
    lbs, ubs = map(np.array, snake.joint_limits())
    tgs = snake.eval_batch(torch.from_numpy(rng.uniform(lbs, ubs, size=(2, 20))).to(dev), want=("ee",))["ee"].contiguous()
    x0s = torch.from_numpy(0.5 * (lbs + ubs)).to(dev)
    cfg4 = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=262144)
    for tile in (0, 0, 0, 32):
        snake.ik_attempts(cfg4, tgs[0], x0s, 262144, tile=tile)
elif which == "single":
    m = np.array(r.fk(rng.uniform(lb, ub))).tolist()
    for _ in range(4):
        r.ik(ob.SolverConfig(), m, list(rng.uniform(lb, ub)))
else:
    T = (1 << 20) if which == "speed" else (1 << 18)
    lb_t, ub_t = torch.from_numpy(lb).to(dev), torch.from_numpy(ub).to(dev)
    g = torch.Generator(device="cuda").manual_seed(42)
    qs = torch.rand((T, 7), dtype=torch.float64, device=dev, generator=g) * (ub_t - lb_t) + lb_t
    x0s = (torch.rand((T, 7), dtype=torch.float64, device=dev, generator=g) * (ub_t - lb_t) + lb_t).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    scfg = ob.SolverConfig(solution_mode=which, max_time=0.0, max_restarts=32)
    for _ in range(3):
        r.ik_batch(scfg, tg, x0s, restarts=32)
torch.cuda.synchronize()
print("done")
