"""Round-2 experiment: thread-per-seed kernel variants (trial columns in local memory = 3 blocks/SM vs shared memory =
2 blocks/SM) on the 65 536-seed step, and Speed-mode batches as dynamic chains vs the static (target, chunk) schedule."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

dev = torch.device("cuda", 0)


def step_times(R=65536, variant=0, reps=10, flush=None, max_evals=0):
    r = ob.Robot.named("panda")
    lb, ub = map(np.array, r.joint_limits())
    rng = np.random.default_rng(42)
    qstar = torch.from_numpy(rng.uniform(lb, ub, size=(reps + 2, 7))).to(dev)
    targets = r.eval_batch(qstar, want=("ee",))["ee"].contiguous()
    x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev)
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    cnt = torch.zeros(3, dtype=torch.int64, device=dev)
    ts = []
    for i in range(reps + 2):
        if flush is not None:
            flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if i == 2: cnt.zero_()
        e0.record()
        out = r.ik_attempts(cfg, targets[i], x0, R, best=True, counters=cnt, variant=variant, max_evals=max_evals)
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    c = cnt.cpu().numpy()
    ms = float(np.median(ts))
    print(f"step R={R} variant={variant} max_evals={max_evals}: median {ms:.4f} ms min {min(ts):.4f}  conv/s={c[2]/reps/ms*1e3:.3e} evals/s={c[1]/reps/ms*1e3:.3e} evals/att={c[1]/c[0]:.2f}", flush=True)


def batch(name, T, R, mode="speed", static=False, variant=0, chunks=0, reps=3):
    r = ob.Robot.named(name)
    n = r.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
    g = torch.Generator(device="cuda").manual_seed(42)
    qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
    cfg = ob.SolverConfig(solution_mode=mode, max_time=0.0, max_restarts=R)
    best = 1e9
    for i in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        q, f, st, ex = r.ik_batch(cfg, tg, x0, restarts=R, chunks=chunks, stats=True, static=static, variant=variant)
        b.record(); torch.cuda.synchronize()
        if i: best = min(best, a.elapsed_time(b))
    cnt = ex["counters"].cpu().numpy()
    ok = float(cfg.is_success(st.cpu().numpy()).mean())
    print(f"{name} T={T} R={R} {mode} static={static} variant={variant}: {best:.3f} ms solves/s={T*ok/best*1e3:.3e} ok={ok:.5f} "
          f"attempts/target={cnt[0]/T:.2f} evals/attempt={cnt[1]/cnt[0]:.2f} evals/s={cnt[1]/best*1e3:.3e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "step"):
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        for v in (1, 2):
            step_times(65536, v, flush=flush)
            step_times(65536, v, flush=flush, max_evals=24)
            step_times(1 << 20, v, reps=3)
    if what in ("all", "batch"):
        for v in (1, 2):
            for T in (1 << 14, 1 << 16, 1 << 18, 1 << 20):
                batch("panda", T, 32, variant=v)
            batch("ur5", 1 << 20, 32, variant=v)
            batch("panda", 1 << 18, 32, mode="quality", variant=v)
        for T in (1 << 14, 1 << 16, 1 << 18, 1 << 20):
            batch("panda", T, 32, static=True)
        batch("ur5", 1 << 20, 32, static=True)
    if what == "tail":
        for T in (1 << 14, 1 << 16):
            for R in (2, 4, 8, 16, 32):
                batch("panda", T, R, variant=2)
        batch("panda", 1 << 16, 32, variant=2, static=True, chunks=1)
