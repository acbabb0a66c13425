"""Quick device-side timing probe (not the bench): attempts/s for each tile width."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

def run(name, T, R, mode, tile, chunks=0, reps=3):
    r = ob.Robot.named(name)
    lb, ub = map(np.array, r.joint_limits())
    rng = np.random.default_rng(42)
    qs = rng.uniform(lb, ub, size=(T, r.num_positions()))
    tg = torch.from_numpy(qs).cuda()
    targets = r.eval_batch(tg, want=("ee",))["ee"]
    x0 = torch.from_numpy(rng.uniform(lb, ub, size=(T, r.num_positions()))).cuda()
    cfg = ob.SolverConfig(solution_mode=mode, max_time=0.0, max_restarts=R)
    best = None
    for i in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        q, f, st, ex = r.ik_batch(cfg, targets, x0, restarts=R, tile=tile, chunks=chunks, stats=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if i > 0: best = ms if best is None else min(best, ms)
    cnt = ex["counters"].cpu().numpy()
    ok = cfg.is_success(st.cpu().numpy()).mean()
    print(f"{name} T={T} R={R} {mode} tile={tile}: {best:.3f} ms  solves/s={T*ok/best*1e3:.3e} success={ok:.4f} "
          f"attempts={cnt[0]} att/s={cnt[0]/best*1e3:.3e} evals/att={cnt[1]/max(cnt[0],1):.1f} evals/s={cnt[1]/best*1e3:.3e}")

if __name__ == "__main__":
    for tile in (1, 8, 32):
        run("panda", 1, 65536, "quality", tile)
    run("panda", 1, 1048576, "quality", 1)
    for tile in (1, 8):
        run("panda", 262144, 32, "speed", tile)
    run("ur5", 1048576, 32, "speed", 1)
    run("snake20", 1, 262144, "quality", 32)
    run("panda", 16384, 256, "quality", 1)
