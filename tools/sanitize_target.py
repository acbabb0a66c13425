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
    if n <= 8:
        # dynamic Speed chains (enough targets to take that path), both column layouts; device path with clamped seeds
        T = 10000
        tgd = r.eval_batch(torch.from_numpy(rng.uniform(lb, ub, size=(T, n))).cuda(), want=("ee",))["ee"].contiguous()
        x0d = torch.from_numpy(rng.uniform(lb, ub, size=(T, n))).cuda()
        x0d[5, 0] = ub[0] + 1.0
        scfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=8)
        for v in (1, 2):
            qd, fd, sd = r.ik_batch(scfg, tgd, x0d, restarts=8, variant=v)
        assert int(sd[5]) & 0x100
        # more targets than resident lanes: exclusive chains that turn shared when idle lanes of their warp join them
        T2 = 40000
        tg2 = r.eval_batch(torch.from_numpy(rng.uniform(lb, ub, size=(T2, n))).cuda(), want=("ee",))["ee"].contiguous()
        x02 = torch.from_numpy(rng.uniform(lb, ub, size=(T2, n))).cuda()
        r.ik_batch(scfg, tg2, x02, restarts=8)
        r.restart_seeds(1, 100)
        r.chacha8_block(np.zeros(8, dtype=np.uint32), 0)
        # Robot::ik fast path (mapped memory, fused selection), default and bounded budgets
        m = [[1, 0, 0, 0.3], [0, 1, 0, 0.1], [0, 0, 1, 0.5], [0, 0, 0, 1]]
        for cfg1 in (ob.SolverConfig(), ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=40)):
            r.ik(cfg1, m, list(x0))
        # peer exchange kernels with three simulated ranks
        lib = ob.load_library()
        W, L = 3, ob.RECORD_HEAD + n
        bufs = [torch.zeros(int(lib.optik_gpu_exchange_bytes(r._h, W)) // 8, dtype=torch.float64, device="cuda") for _ in range(W)]
        peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
        st_ = torch.cuda.current_stream().cuda_stream
        for w in range(W):
            ob._check(lib.optik_gpu_exchange_push(r._h, rec[w].data_ptr(), peers.data_ptr(), w, W, 1, st_))
        out_ = torch.empty(L, dtype=torch.float64, device="cuda")
        ob._check(lib.optik_gpu_exchange_select(r._h, bufs[0].data_ptr(), W, 1, out_.data_ptr(), st_))
        # fused selection + fused push inside the per-attempt launch
        cfgq = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=300)
        r.ik_attempts(cfgq, tgd[0], x0d[0], 300, best=True, push=(peers.data_ptr(), 1, W, 2))
    torch.cuda.synchronize()
print("sanitize target done")
