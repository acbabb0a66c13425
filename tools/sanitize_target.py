"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

rng = np.random.default_rng(0)
for name in ("panda", "ur5", "snake20"):
    r = ob.Robot.named(name)
    lb, ub = map(np.array, r.joint_limits())
    n = len(lb)
    B = 301
    q = rng.uniform(lb, ub, size=(B, n))
    tg = r.eval_batch(rng.uniform(lb, ub, size=(B, n)), want=("ee",))["ee"]
    out = r.eval_batch(q, tg)
    out1 = r.eval_batch(q[1:], tg[1:])  # unaligned q (odd n): plain-load tile path
    assert np.array_equal(out["f"][1:], out1["f"])
    x0 = 0.5 * (lb + ub)
    for tile in ((1, 8, 32) if n <= 8 else (32,)):
        cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=200)
        r.ik_attempts(cfg, tg[0], x0, 200, tile=tile, best=True)
        scfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=6)
        r.ik_batch(scfg, tg[:70], np.tile(x0, (70, 1)), restarts=6, tile=tile)
        r.ik_batch(scfg, tg[:3], np.tile(x0, (3, 1)), restarts=6, tile=tile, chunks=3)
    if n in (6, 7):
        a, v, st = r.diff_ik_batch(q, rng.random((B, 6)), np.ones(n))
        assert st.all()
    rec = torch.rand((5, ob.RECORD_HEAD + n), dtype=torch.float64, device="cuda")
    r.select_records(rec)
    torch.cuda.synchronize()
print("sanitize target done")
