set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python - <<'PY'
import numpy as np, torch, optik_b200 as ob
dev = torch.device("cuda", 0)
snake = ob.Robot.named("snake20")
lbs, ubs = map(np.array, snake.joint_limits())
R4 = 262144
cfg4 = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R4)
tgs = snake.eval_batch(torch.from_numpy(np.random.default_rng(42).uniform(lbs, ubs, size=(4, 20))).to(dev), want=("ee",))["ee"].contiguous()
x0s = torch.from_numpy(0.5 * (lbs + ubs)).to(dev)
for tile in (0, 32):
    best = 1e9
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); q, f, st, ev = snake.ik_attempts(cfg4, tgs[i], x0s, R4, tile=tile); b.record(); torch.cuda.synchronize()
        if i: best = min(best, a.elapsed_time(b))
    print(f"snake20 tile={tile}: {best:.3f} ms conv/s={(st==1).sum().item()/best*1e3:.3e} evals/att={ev.double().mean().item():.2f}")
PY
