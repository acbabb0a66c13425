"""GPU: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars (north star): evaluator vs the golden-pinned fp64 oracle <= 1e-12 relative (fp64 arithmetic, different but
equivalent formulas); solver vs its fp64 CPU twin: identical status / evaluation count per restart seed and
solutions within 1e-6 rad (in practice bit-identical, which is asserted separately so a regression is visible).
"""
import numpy as np
import pytest

import optik_b200 as ob
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RAD_TOL = 1e-6  # north star: <= 1e-6 rad on converged seeds


def robot_and_chain(name):
    r = ob.Robot.named(name)
    return r, O.Chain(r.chain())


def twin_layout(ch, tile=0):
    """Which twin mirrors the kernel the library picks: thread-per-seed (1) for tile=1 or auto in batched calls (every
    chain: long ones run 32-thread blocks so that their rows fit shared memory), else the tile kernel's
    lane-per-joint order (0)."""
    return 1 if tile in (0, 1) else 0


def targets_for(ch, rng, T):
    out = np.zeros((T, 8))
    for t in range(T):
        out[t] = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
    return out


# ------------------------------------------------------------------ evaluator
def test_fk_against_reference_golden(golden):  # tests/test_fk.rs:13-26 through the CUDA evaluator
    r, _ = robot_and_chain("ur3e")
    q = np.array(golden["fk_inputs"])
    ee = r.eval_batch(q, want=("ee",))["ee"]
    ref = np.array(golden["fk_outputs"])
    assert np.abs(ee[:, 4:7] - ref[:, 4:7]).max() < 1e-12
    sign = np.sign(np.sum(ee[:, :4] * ref[:, :4], axis=1))[:, None]
    assert np.abs(ee[:, :4] * sign - ref[:, :4]).max() < 1e-12


def test_jacobian_objective_gradient_against_pinned_oracle(golden):
    r, _ = robot_and_chain("ur3e")
    q = np.array(golden["fk_inputs"])
    wl, wa = golden["oracle_weights"]
    out = r.eval_batch(q, np.array(golden["oracle_target"]), wl, wa)
    assert np.abs(out["jac"] - np.array(golden["oracle_jacobian"])).max() < 1e-12
    f_ref, g_ref = np.array(golden["oracle_objective"]), np.array(golden["oracle_gradient"])
    assert np.abs(out["f"] - f_ref).max() <= 1e-12 * np.abs(f_ref).max()
    assert np.abs(out["grad"] - g_ref).max() <= 1e-12 * np.abs(g_ref).max()


@pytest.mark.parametrize("name", ["panda", "ur5", "snake20"])
def test_evaluator_vs_oracle_random(name):
    r, ch = robot_and_chain(name)
    rng = np.random.default_rng(2)
    B = 257  # ragged: not a multiple of the block size
    q = rng.uniform(ch.lb, ch.ub, size=(B, ch.n))
    tg = targets_for(ch, rng, B)
    out = r.eval_batch(q, tg)
    for i in range(0, B, 8):
        _, ee = ch.fk(q[i])
        assert np.abs(out["ee"][i, :7] - ee[:7]).max() < 1e-12
        assert np.abs(out["jac"][i].reshape(ch.n, 6).T - ch.joint_jacobian(q[i])).max() < 1e-12
        f, g = ch.objective(q[i], tg[i]), ch.objective_grad(q[i], tg[i])
        assert abs(out["f"][i] - f) <= 1e-12 * max(1.0, f)
        assert np.abs(out["grad"][i] - g).max() <= 1e-11 * max(1.0, np.abs(g).max())


def test_evaluator_ee_offset_and_single_calls():
    r, ch = robot_and_chain("ur3e")
    q = np.array([0.3, -1.0, 0.8, 0.1, 1.2, -0.4])
    off = O.pose8([0.1, 0.2, -0.1, 0.9695], [0.05, -0.02, 0.1])
    off[:4] /= np.linalg.norm(off[:4])
    M = np.array(r.fk(q, ee_offset=O.pose8_to_matrix(off)))
    _, ee = ch.fk(q, off)
    assert np.abs(M - O.pose8_to_matrix(ee)).max() < 1e-12
    J = np.array(r.joint_jacobian(q, ee_offset=O.pose8_to_matrix(off)))
    assert np.abs(J - ch.joint_jacobian(q, off)).max() < 1e-12
    # C ABI single calls: column-major 4x4 / 6xn, caller frees (crates/optik-cpp/src/lib.rs:90-116)
    import ctypes as C
    lib = ob.load_library()
    x = (C.c_double * 6)(*q)
    m = ob._take(lib.optik_robot_fk(r._h, x), 16).reshape(4, 4).T
    assert np.abs(m - O.pose8_to_matrix(ch.fk(q)[1])).max() < 1e-12
    jac = ob._take(lib.optik_robot_joint_jacobian(r._h, x), 36).reshape(6, 6).T
    assert np.abs(jac - ch.joint_jacobian(q)).max() < 1e-12


# ------------------------------------------------------------------ solver vs twin, per restart seed
@pytest.mark.parametrize("name,tile,R", [("panda", 1, 4096), ("ur5", 1, 2048), ("ur3e", 1, 1024),
                                          ("panda", 8, 2048), ("panda", 32, 512), ("ur5", 8, 1024), ("ur5", 16, 256),
                                          ("ur3e", 8, 512), ("snake20", 32, 512), ("snake20", 0, 256), ("panda", 0, 512)])
def test_attempts_match_twin_per_seed(name, tile, R):
    r, ch = robot_and_chain(name)
    rng = np.random.default_rng(42)
    tgt = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
    x0 = 0.5 * (ch.lb + ch.ub)
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    q, f, st, ev = r.ik_attempts(cfg, tgt, x0, R, tile=tile)
    tq, tf, tst, tev = O.twin_attempts(ch, tgt, x0, 0, R, O.twin_params(layout=twin_layout(ch, tile)))
    assert np.array_equal(st, tst), "success set / status differs from the CPU twin"
    assert np.array_equal(ev, tev)
    ok = st == 1
    assert ok.sum() > R // 10
    assert np.abs(q[ok] - tq[ok]).max() <= RAD_TOL
    # stronger, informational-but-asserted: the arithmetic spec makes the two bit-identical
    assert np.array_equal(q, tq) and np.array_equal(f, tf)
    # every converged seed passes the reference's success predicate under the golden-pinned oracle
    for i in np.where(ok)[0][:64]:
        assert ch.objective(q[i], tgt) < cfg.tol_f
        assert np.all(q[i] >= ch.lb) and np.all(q[i] <= ch.ub)


def test_attempts_with_weights_tolerances_and_offsets():
    r, ch = robot_and_chain("ur3e")
    rng = np.random.default_rng(9)
    x0 = np.zeros(6)
    off = O.pose8([0, 0, 0.3826834, 0.9238795], [0.0, 0.05, 0.1])
    off[:4] /= np.linalg.norm(off[:4])
    tgt = ch.fk(rng.uniform(ch.lb, ch.ub), off)[1]
    for kw in (dict(tol_f=1e-12), dict(tol_df=1e-4), dict(tol_dx=1e-3), dict(linear_weight=[1.0, 2.0, 0.5], angular_weight=[0.1, 1.0, 1.0])):
        cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=128, **kw)
        for tile in (1, 8):
            P = O.twin_params(tol_f=cfg.tol_f, tol_df=cfg.tol_df, tol_dx=cfg.tol_dx, wl=cfg.linear_weight,
                              wa=cfg.angular_weight, layout=twin_layout(ch, tile))
            q, f, st, ev = r.ik_attempts(cfg, tgt, x0, 128, restart_begin=5, ee_offset=off, tile=tile)
            tq, tf, tst, tev = O.twin_attempts(ch, tgt, x0, 5, 133, P, ee_offset=off)
            assert np.array_equal(st, tst) and np.array_equal(ev, tev), (kw, tile)
            assert np.abs(q - tq).max() <= RAD_TOL, (kw, tile)
            assert cfg.is_success(st).sum() > 5, (kw, tile)


# ------------------------------------------------------------------ batched ik(): selection semantics
@pytest.mark.parametrize("mode", ["speed", "quality"])
@pytest.mark.parametrize("chunks", [1, 4, 0])
@pytest.mark.parametrize("tile", [0, 8])
def test_ik_batch_matches_reference_selection(mode, chunks, tile):
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(7)
    T, R = 96, 12
    tg = targets_for(ch, rng, T)
    x0 = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
    cfg = ob.SolverConfig(solution_mode=mode, max_time=0.0, max_restarts=R)
    q, f, st, extra = r.ik_batch(cfg, tg, x0, restarts=R, chunks=chunks, tile=tile, stats=True)
    P = O.twin_params(layout=twin_layout(ch, tile))
    for t in range(T):
        ref = O.twin_ik(ch, tg[t], x0[t], 0, R, mode, P)
        assert bool(cfg.is_success(st[t])) == ref["found"], t
        if ref["found"]:
            assert int(extra["restart"][t]) == ref["restart"], t
            assert np.abs(q[t] - ref["q"]).max() <= RAD_TOL
            assert ch.objective(q[t], tg[t]) < cfg.tol_f
    assert cfg.is_success(st).mean() > 0.9


def test_ik_batch_device_path_matches_host_path():
    import torch
    r, ch = robot_and_chain("ur5")
    rng = np.random.default_rng(3)
    T = 1000
    tg = targets_for(ch, rng, T)
    x0 = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=8)
    q, f, st = r.ik_batch(cfg, tg, x0, restarts=8)
    dq, df, dst = r.ik_batch(cfg, torch.from_numpy(tg).cuda(), torch.from_numpy(x0).cuda(), restarts=8)
    torch.cuda.synchronize()
    assert np.array_equal(dq.cpu().numpy(), q) and np.array_equal(dst.cpu().numpy(), st)


def test_restart_sharding_is_consistent():
    """Splitting the restart range (multi-GPU restart sharding, waves) and re-selecting gives the same answer."""
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(5)
    T = 32
    tg = targets_for(ch, rng, T)
    x0 = np.tile(0.5 * (ch.lb + ch.ub), (T, 1))
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=16)
    q, f, st = r.ik_batch(cfg, tg, x0, restarts=16)
    qa, fa, sa = r.ik_batch(cfg, tg, x0, restarts=8, restart_begin=0)
    qb, fb, sb = r.ik_batch(cfg, tg, x0, restarts=8, restart_begin=8)
    for t in range(T):
        cands = [(np.sum((qq[t] - x0[t]) ** 2), qq[t]) for qq, ss in ((qa, sa), (qb, sb)) if cfg.is_success(ss[t])]
        assert bool(cands) == bool(cfg.is_success(st[t]))
        if cands:
            assert np.array_equal(min(cands, key=lambda c: c[0])[1], q[t])


# ------------------------------------------------------------------ the reference's own ik() tests (tests/test_ik.rs)
def test_solution_forward_backward():  # tests/test_ik.rs:91-130
    r, ch = robot_and_chain("ur3e")
    rng = np.random.default_rng(42)
    cfg = ob.SolverConfig(solution_mode="speed", tol_f=1e-12, max_time=0.0, max_restarts=25)
    for _ in range(10):
        x_target = rng.random(6)
        T = np.array(r.fk(x_target))
        sol = r.ik(cfg, T, [0.0] * 6)
        assert sol is not None
        assert np.abs(np.array(r.fk(sol[0])) - T).max() < 1e-6


def test_determinism():  # tests/test_ik.rs:45-89
    r, _ = robot_and_chain("ur3e")
    r.set_parallelism(1)
    T = np.array(r.fk(np.random.default_rng(42).random(6)))
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=25)
    first = r.ik(cfg, T, [0.0] * 6)
    assert first is not None
    for _ in range(10):
        again = r.ik(cfg, T, [0.0] * 6)
        assert np.abs(np.array(again[0]) - np.array(first[0])).max() < 1e-6


def test_solution_quality():  # tests/test_ik.rs:132-182
    r, _ = robot_and_chain("ur3e")
    rng = np.random.default_rng(42)
    speed = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=15)
    quality = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=15)
    for _ in range(20):
        T = np.array(r.fk(rng.random(6)))
        x0 = [0.0] * 6
        s, q = r.ik(speed, T, x0), r.ik(quality, T, x0)
        assert s is not None and q is not None
        assert np.linalg.norm(np.array(q[0]) - x0) <= np.linalg.norm(np.array(s[0]) - x0)


def test_stopping_maxtime():  # tests/test_ik.rs:24-43: impossible goal returns None after ~max_time
    import time
    r, _ = robot_and_chain("ur3e")
    T = np.eye(4)
    T[:3, 3] = 100.0
    r.ik(ob.SolverConfig(max_time=0.05), T, [0.0] * 6)  # warm-up (context, module load)
    t0 = time.perf_counter()
    out = r.ik(ob.SolverConfig(max_time=0.05), T, [0.0] * 6)
    dt = time.perf_counter() - t0
    assert out is None
    assert abs(dt - 0.05) < 0.1


def test_c_abi_ik_returns_malloced_solution():  # crates/optik-cpp/src/lib.rs:127-162, lib.cpp:105-120
    import ctypes as C
    r, ch = robot_and_chain("ur3e")
    lib = ob.load_library()
    qstar = np.array([0.5, -0.7, 0.9, 0.2, -0.4, 0.3])
    M = O.pose8_to_matrix(ch.fk(qstar)[1])
    tgt = (C.c_double * 16)(*M.T.ravel())  # column-major
    x0 = (C.c_double * 6)(*([0.0] * 6))
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=25)._c()
    p = lib.optik_robot_ik(r._h, C.byref(cfg), tgt, x0)
    assert p
    q = ob._take(p, 6)
    assert ch.objective(q, ch.fk(qstar)[1]) < 1e-6
    far = np.eye(4)
    far[:3, 3] = 100.0
    tgt2 = (C.c_double * 16)(*far.T.ravel())
    assert not lib.optik_robot_ik(r._h, C.byref(cfg), tgt2, x0)  # NULL == no solution


# ------------------------------------------------------------------ full-size properties (BASELINE configs)
def test_config2_full_size_properties():
    """Panda, 65 536 restart seeds to one target: every converged record re-evaluates below tol_f inside the limits
    (checked on the GPU evaluator, spot-checked by the oracle), and a sample of seeds matches the twin exactly."""
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(42)
    tgt = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
    x0 = 0.5 * (ch.lb + ch.ub)
    R = 65536
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    q, f, st, ev = r.ik_attempts(cfg, tgt, x0, R)
    ok = st == 1
    assert 0.2 < ok.mean() < 0.95
    fe = r.eval_batch(q, tgt, want=("f",))["f"]
    assert np.all(fe[ok] < cfg.tol_f)
    assert np.allclose(fe, f, rtol=1e-9, atol=1e-18)  # evaluator kernel (backward recursion) vs solve kernel (scan)
    assert np.all(q[ok] >= ch.lb) and np.all(q[ok] <= ch.ub)
    idx = rng.choice(R, 256, replace=False)
    for i in idx:
        tq, tf, tst, tev = O.twin_attempts(ch, tgt, x0, int(i), int(i) + 1, O.twin_params(layout=twin_layout(ch)))
        assert tst[0] == st[i] and tev[0] == ev[i] and np.array_equal(tq[0], q[i])
    # Quality selection over all seeds == arg-min distance among converged
    qb, fb, sb = r.ik_batch(cfg, tgt[None, :], x0[None, :], restarts=R)
    d = np.sum((q[ok] - x0) ** 2, axis=1)
    assert np.array_equal(qb[0], q[ok][np.argmin(d)])


def _device_targets(r, ch, T, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    lb, ub = torch.from_numpy(ch.lb).cuda(), torch.from_numpy(ch.ub).cuda()
    qs = torch.rand((T, ch.n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, ch.n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    return r.eval_batch(qs, want=("ee",))["ee"].contiguous(), x0, lb, ub


def test_config3_full_size_properties():
    """UR5, 1 Mi independent reachable targets, Speed, <= 32 restarts (BASELINE config 3): every target reported solved
    re-evaluates below tol_f inside the limits; the winner is the LOWEST converged restart index (= the reference with
    one thread, lib.rs:409-412) on a sample checked against the twin; per-target success ~ 1."""
    import torch
    r, ch = robot_and_chain("ur5")
    T, R = 1 << 20, 32
    tg, x0, lb, ub = _device_targets(r, ch, T, 42)
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q, f, st, extra = r.ik_batch(cfg, tg, x0, restarts=R, stats=True)
    ok = torch.as_tensor(cfg.is_success(st.cpu().numpy()), device="cuda")
    assert float(ok.double().mean()) > 0.9995
    fe = r.eval_batch(q, tg, want=("f",))["f"]
    assert bool((fe[ok] < cfg.tol_f).all()) and bool(((q[ok] >= lb) & (q[ok] <= ub)).all())
    assert int(extra["restart"][ok].max()) < R
    idx = torch.randperm(T, generator=torch.Generator().manual_seed(1))[:48]
    tg_h, x0_h, q_h, rs_h = tg[idx].cpu().numpy(), x0[idx].cpu().numpy(), q[idx].cpu().numpy(), extra["restart"][idx].cpu().numpy()
    P = O.twin_params(layout=twin_layout(ch))
    for k in range(len(idx)):
        ref = O.twin_ik(ch, tg_h[k], x0_h[k], 0, R, "speed", P)
        assert ref["found"] and int(rs_h[k]) == ref["restart"] and np.array_equal(q_h[k], ref["q"])


def test_config4_full_size_properties():
    """20-DOF snake, 262 144 seeds to one target (BASELINE config 4): converged records re-evaluate below tol_f inside
    the (tight) limits; a sample of seeds is bit-identical to the twin; pinned joints sit ON a limit.  Runs the default
    layout (thread-per-seed kernel, one column row per thread) and the tile kernel (one warp per seed)."""
    r, ch = robot_and_chain("snake20")
    rng = np.random.default_rng(42)
    tgt = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
    x0 = 0.5 * (ch.lb + ch.ub)
    R = 262144
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    for tile in (0, 32):
        q, f, st, ev = r.ik_attempts(cfg, tgt, x0, R if tile == 0 else R // 8, tile=tile)
        ok = st == 1
        assert ok.mean() > 0.8
        fe = r.eval_batch(q, tgt, want=("f",))["f"]
        assert np.all(fe[ok] < cfg.tol_f) and np.all(q >= ch.lb) and np.all(q <= ch.ub)
        assert ((q[ok] == ch.lb) | (q[ok] == ch.ub)).any()  # the joint-limit stress really exercises the projection
        for i in rng.choice(len(st), 64, replace=False):
            tq, tf, tst, tev = O.twin_attempts(ch, tgt, x0, int(i), int(i) + 1, O.twin_params(layout=twin_layout(ch, tile)))
            assert tst[0] == st[i] and tev[0] == ev[i] and np.array_equal(tq[0], q[i])


def test_config5_shard_full_size_properties():
    """Panda, Quality, 256 restarts x 131 072 targets = one GPU's shard of BASELINE config 5: every target solved and
    re-verified; Quality is never farther from the seed than Speed on the same targets (tests/test_ik.rs:132-182 at
    scale); a sample equals the twin's arg-min over all 256 restarts."""
    import torch
    r, ch = robot_and_chain("panda")
    T, R = 131072, 256
    tg, x0, lb, ub = _device_targets(r, ch, T, 7)
    qual = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    speed = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q, f, st = r.ik_batch(qual, tg, x0, restarts=R)
    qs, fs, ss = r.ik_batch(speed, tg, x0, restarts=R)
    ok = torch.as_tensor(qual.is_success(st.cpu().numpy()), device="cuda")
    oks = torch.as_tensor(speed.is_success(ss.cpu().numpy()), device="cuda")
    assert float(ok.double().mean()) > 0.9999 and bool((ok == oks).all())
    fe = r.eval_batch(q, tg, want=("f",))["f"]
    assert bool((fe[ok] < qual.tol_f).all()) and bool(((q[ok] >= lb) & (q[ok] <= ub)).all())
    dq, ds = ((q - x0) ** 2).sum(dim=1), ((qs - x0) ** 2).sum(dim=1)
    assert bool((dq[ok] <= ds[ok]).all())
    idx = torch.randperm(T, generator=torch.Generator().manual_seed(2))[:6]
    P = O.twin_params(layout=twin_layout(ch))
    for k in idx.tolist():
        ref = O.twin_ik(ch, tg[k].cpu().numpy(), x0[k].cpu().numpy(), 0, R, "quality", P)
        assert ref["found"] and np.array_equal(q[k].cpu().numpy(), ref["q"])


# ------------------------------------------------------------------ best-pick records (cross-GPU exchange building blocks)
@pytest.mark.parametrize("mode", ["speed", "quality"])
def test_attempts_best_record_matches_reference_selection(mode):
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(21)
    tgt = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
    x0 = rng.uniform(ch.lb, ch.ub)
    cfg = ob.SolverConfig(solution_mode=mode, max_time=0.0, max_restarts=300)
    q, f, st, ev, rec = r.ik_attempts(cfg, tgt, x0, 300, restart_begin=40, best=True)
    ref = O.twin_ik(ch, tgt, x0, 40, 340, mode, O.twin_params(layout=twin_layout(ch)))
    assert ref["found"] and rec[0] == 1.0
    assert int(rec[2]) == ref["restart"] and rec[4] == ref["status"] and rec[3] == ref["f"]
    assert np.array_equal(rec[8:], ref["q"])
    if mode == "quality":
        assert np.isclose(rec[1], np.sum((ref["q"] - x0) ** 2), rtol=1e-12)
    else:
        assert rec[1] == ref["restart"]


def test_select_records_kernel_matches_rule():
    import torch
    from optik_b200 import dist as obd
    r, _ = robot_and_chain("panda")
    rng = np.random.default_rng(8)
    for count in (1, 2, 8, 33, 100):
        rec = rng.random((count, obd.RECORD_HEAD + 7))
        rec[:, 0] = rng.random(count) < 0.6
        rec[:, 1] = np.round(rec[:, 1], 1)  # force score ties
        rec[:, 2] = rng.permutation(count)
        t = torch.from_numpy(rec).cuda()
        got = r.select_records(t).cpu()
        idx, want = obd.select_candidates(torch.from_numpy(rec))
        if rec[:, 0].max() > 0:
            assert torch.equal(got, want), count
        else:
            assert got[0] == 0.0


# ------------------------------------------------------------------ chain shapes beyond the bundled robots
def synth_urdf(joints, tip=None):
    """joints: list of (type, xyz, rpy, axis, lo, hi)."""
    s = ['<robot name="synth">', '<link name="l0"/>']
    for i, (typ, xyz, rpy, axis, lo, hi) in enumerate(joints):
        s.append(f'<link name="l{i+1}"/>')
        s.append(f'<joint name="j{i}" type="{typ}"><parent link="l{i}"/><child link="l{i+1}"/>'
                 f'<origin xyz="{xyz}" rpy="{rpy}"/><axis xyz="{axis}"/>'
                 f'<limit lower="{lo}" upper="{hi}" effort="1" velocity="1"/></joint>')
    last = f"l{len(joints)}"
    if tip:
        s.append('<link name="tip"/>')
        s.append(f'<joint name="jt" type="fixed"><parent link="{last}"/><child link="tip"/><origin xyz="{tip[0]}" rpy="{tip[1]}"/></joint>')
        last = "tip"
    s.append("</robot>")
    return "\n".join(s), "l0", last


def arm(n, prismatic_at=(), tip=("0 0.05 0.1", "0.3 0 0.2")):
    rng = np.random.default_rng(n)
    js = []
    for i in range(n):
        axis = ["0 0 1", "0 1 0", "1 0 0", "0.6 0 0.8"][i % 4]
        xyz = " ".join(f"{v:.3f}" for v in rng.uniform(-0.15, 0.25, 3))
        rpy = " ".join(f"{v:.3f}" for v in rng.uniform(-1.0, 1.0, 3))
        if i in prismatic_at:
            js.append(("prismatic", xyz, rpy, axis, -0.2, 0.3))
        else:
            js.append(("revolute", xyz, rpy, axis, -2.5, 2.5))
    return synth_urdf(js, tip)


@pytest.mark.parametrize("n,pris,tip,tiles", [
    (8, (), ("0 0.05 0.1", "0.3 0 0.2"), (1, 8, 16)),      # n == tile width: the tip cannot ride the scan (8 lanes)
    (8, (), None, (1, 8)),                                   # no tip joint at all
    (16, (), ("0 0 0.1", "0 0 0"), (1, 16, 32)),             # n == 16
    (5, (2,), ("0.1 0 0", "0 0.5 0"), (1, 8)),               # prismatic joint mid-chain (reference FK supports it)
    (4, (0, 3), None, (1, 8, 32)),                           # prismatic first and last
    (1, (), ("0 0 0.2", "0 0 0"), (1, 8)),                   # single joint
    (12, (5,), ("0 0 0.1", "0.1 0 0"), (1, 16, 32)),         # n > 8 (seeds span two ChaCha8 blocks), a prismatic joint
    (27, (), ("0 0 0.05", "0 0 0"), (1, 32)),                # long chains: 32-thread blocks of the thread-per-seed kernel
    (32, (), ("0 0 0.05", "0 0 0"), (0, 1, 32)),             # the longest chain the library takes (auto = thread per seed)
])
def test_chain_shapes_match_twin_and_oracle(n, pris, tip, tiles):
    urdf, base, ee = arm(n, pris, tip)
    r = ob.Robot.from_urdf_str(urdf, base, ee)
    ch = O.Chain(r.chain())
    assert ch.n == n
    rng = np.random.default_rng(100 + n)
    # evaluator vs the golden-pinned oracle (FK + Jacobian incl. prismatic columns + objective + gradient)
    q = rng.uniform(ch.lb, ch.ub, size=(33, n))
    tg = targets_for(ch, rng, 33)
    out = r.eval_batch(q, tg)
    for i in range(0, 33, 4):
        assert np.abs(out["ee"][i, :7] - ch.fk(q[i])[1][:7]).max() < 1e-12
        assert np.abs(out["jac"][i].reshape(n, 6).T - ch.joint_jacobian(q[i])).max() < 1e-12
        assert abs(out["f"][i] - ch.objective(q[i], tg[i])) <= 1e-12 * max(1.0, out["f"][i])
        g = ch.objective_grad(q[i], tg[i])
        assert np.abs(out["grad"][i] - g).max() <= 1e-11 * max(1.0, np.abs(g).max())
    # solver vs twin per seed, every layout that fits
    tgt = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
    x0 = 0.5 * (ch.lb + ch.ub)
    R = 256
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    for tile in tiles:
        qq, f, st, ev = r.ik_attempts(cfg, tgt, x0, R, tile=tile)
        tq, tf, tst, tev = O.twin_attempts(ch, tgt, x0, 0, R, O.twin_params(layout=twin_layout(ch, tile)))
        assert np.array_equal(st, tst) and np.array_equal(ev, tev), tile
        assert np.array_equal(qq, tq), tile
        ok = st == 1
        if n > 1:
            assert ok.sum() > 0
        for i in np.where(ok)[0][:16]:
            assert ch.objective(qq[i], tgt) < cfg.tol_f and np.all(qq[i] >= ch.lb) and np.all(qq[i] <= ch.ub)


def test_edge_cases_empty_ragged_and_limits():
    r, ch = robot_and_chain("ur5")
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=4)
    # empty batch
    q, f, st = r.ik_batch(cfg, np.zeros((0, 8)), np.zeros((0, 6)), restarts=4)
    assert q.shape == (0, 6) and st.shape == (0,)
    assert r.eval_batch(np.zeros((0, 6)))["ee"].shape == (0, 8)
    # T = 1, R = 1 (restart 0 only = the caller's seed), seed exactly on a solution -> converges at evaluation 1
    qs = np.array([0.3, -0.4, 0.5, 0.1, -0.2, 0.6])
    tgt = ch.fk(qs)[1]
    q, f, st, extra = r.ik_batch(cfg, tgt[None], qs[None], restarts=1, stats=True)
    assert st[0] == 1 and extra["evals"][0] == 1 and np.array_equal(q[0], qs)
    # seeds outside the limits are rejected on the host path (lib.rs:251-254)
    bad = qs.copy()
    bad[2] = ch.ub[2] + 1.0
    with pytest.raises(ob.OptikError, match="joint limits"):
        r.ik_batch(cfg, tgt[None], bad[None], restarts=4)
    # seed exactly on the limits is legal
    edge = ch.ub.copy()
    q, f, st = r.ik_batch(cfg, tgt[None], edge[None], restarts=4)
    assert st.shape == (1,)
    # more chunks than restarts, odd sizes
    T = 37
    rng = np.random.default_rng(4)
    tg = targets_for(ch, rng, T)
    x0 = rng.uniform(ch.lb, ch.ub, size=(T, 6))
    a = r.ik_batch(cfg, tg, x0, restarts=3, chunks=7)
    b = r.ik_batch(cfg, tg, x0, restarts=3, chunks=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_infinite_limits_chain_solves():
    """upper-lower <= 0 means unlimited (kinematics.rs:299-303); restart seeds then draw from [-pi, pi] (documented)."""
    js = [("revolute", "0 0 0.1", "0 0 0", "0 0 1", 0, 0), ("revolute", "0.2 0 0", "1.2 0 0", "0 0 1", 0, 0),
          ("revolute", "0.2 0 0", "0 0.7 0", "0 1 0", -2, 2), ("revolute", "0.1 0 0.1", "0 0 0.4", "1 0 0", 0, 0),
          ("revolute", "0 0.1 0", "0.5 0 0", "0 0 1", -2, 2), ("revolute", "0 0 0.1", "0 0.9 0", "0 1 0", 0, 0)]
    urdf, base, ee = synth_urdf(js, ("0 0 0.1", "0 0 0"))
    r = ob.Robot.from_urdf_str(urdf, base, ee)
    ch = O.Chain(r.chain())
    assert np.isinf(ch.lb[0]) and np.isinf(ch.ub[0])
    tgt = ch.fk(np.array([0.4, -0.3, 0.8, 0.2, -0.5, 0.1]))[1]
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=128)
    q, f, st, ev = r.ik_attempts(cfg, tgt, np.zeros(6), 128)
    tq, tf, tst, tev = O.twin_attempts(ch, tgt, np.zeros(6), 0, 128, O.twin_params(layout=1))
    assert np.array_equal(st, tst) and np.array_equal(q, tq)
    assert (st == 1).sum() > 10
    for i in np.where(st == 1)[0][:8]:
        assert ch.objective(q[i], tgt) < 1e-6


def test_async_host_calls_pipeline_and_match_sync():
    """OPTIK_BATCH_ASYNC: two streams, two sets of pinned buffers, calls enqueued back to back -- same records as the
    synchronous call, bit for bit."""
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(9)
    R, n = 2048, ch.n
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    x0 = ob.pinned_empty(n)
    x0[:] = 0.5 * (ch.lb + ch.ub)
    tgts = ob.pinned_empty((4, 8))
    tgts[:] = targets_for(ch, rng, 4)
    streams = [ob.Stream(r), ob.Stream(r)]
    bufs = [(ob.pinned_empty((R, n)), ob.pinned_empty(R), ob.pinned_empty(R, np.int32), ob.pinned_empty(R, np.int32))
            for _ in range(2)]
    recs = [ob.pinned_empty(ob.RECORD_HEAD + n) for _ in range(2)]
    got = []
    for s in range(4):
        k = s & 1
        if s >= 2:
            streams[k].synchronize()
            got.append(tuple(a.copy() for a in bufs[k]) + (recs[k].copy(),))
        r.ik_attempts(cfg, tgts[s], x0, R, best=True, out=bufs[k], record=recs[k], stream=streams[k], wait=False)
    for k in (0, 1):
        streams[k].synchronize()
        got.append(tuple(a.copy() for a in bufs[k]) + (recs[k].copy(),))
    for s in range(4):
        q, f, st, ev, best = r.ik_attempts(cfg, np.array(tgts[s]), np.array(x0), R, best=True)
        for a, b in zip(got[s], (q, f, st, ev, best)):
            assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        r.ik_attempts(cfg, tgts[0], x0, R, out=bufs[0], wait=False)  # no stream


# ------------------------------------------------------------------ diff_ik (lib.rs:101-239)
def test_diff_ik_matches_lp_golden_and_oracle():
    """CUDA diff_ik vs the HiGHS-solved LP fixtures (alpha unique; v unique for 6 DOF) and vs the oracle's exact
    solution on the same inputs (tolerance 1e-9: fp64, different but equivalent FK formulas)."""
    import json, os
    from conftest import ROOT
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "diff_ik_vectors.json")))["cases"]
    robots = {name: robot_and_chain(name) for name in ("ur3e", "panda")}
    for c in cases:
        r, ch = robots[c["robot"]]
        res = r.diff_ik(c["x0"], c["V_WE"], c["v_max"])
        assert res is not None
        alpha, v = res
        assert abs(alpha - c["alpha"]) <= 1e-8
        if c.get("singular"):
            continue  # alpha is the LP's unique optimum; v is not (any null-space vector inside the box)
        if c["v"] is not None:
            assert np.abs(np.array(v) - np.array(c["v"])).max() <= 1e-8
        oa, ov = ch.diff_ik(c["x0"], c["V_WE"], c["v_max"])
        assert abs(alpha - oa) <= 1e-9 * max(1.0, abs(oa)) and np.abs(np.array(v) - ov).max() <= 1e-9 * max(1.0, np.abs(ov).max())


@pytest.mark.parametrize("name", ["ur3e", "ur5", "panda"])
def test_diff_ik_batch_properties(name):
    """tests/test_ik.rs:184-209 at batch size (ragged B): alpha in [0,1], |v| <= vmax, J_W v = alpha V (the reference's
    TODO), checked for every configuration; a sample against the oracle; shared and per-row V / vmax agree."""
    import torch
    r, ch = robot_and_chain(name)
    rng = np.random.default_rng(7)
    B, n = 4099, ch.n
    x0 = rng.uniform(ch.lb, ch.ub, size=(B, n))
    V = rng.random((B, 6))
    vmax = rng.uniform(0.2, 2.0, size=(B, n))
    alpha, v, st = r.diff_ik_batch(x0, V, vmax)
    ok = st == 1
    assert ok.mean() > 0.999
    assert np.all(alpha[ok] >= 0) and np.all(alpha[ok] <= 1 + 1e-12)
    assert np.all(np.abs(v[ok]) <= vmax[ok] * (1 + 1e-12))
    out = r.eval_batch(x0, want=("ee", "jac"))
    Jb = out["jac"].reshape(B, n, 6)  # column-major 6 x n per row
    ee = out["ee"]
    def rot(q, u):
        w, uv = q[:, 3:4], q[:, :3]
        t = 2.0 * np.cross(uv, u)
        return u + w * t + np.cross(uv, t)
    tw = np.einsum("bnk,bn->bk", Jb, v)  # body-frame twist J v
    tw_w = np.concatenate([rot(ee, tw[:, :3]), rot(ee, tw[:, 3:])], axis=1)
    resid = np.abs(tw_w - alpha[:, None] * V)[ok].max(axis=1)
    assert np.quantile(resid, 0.99) < 1e-9 and resid.max() < 1e-6  # near-singular rows amplify rounding
    for i in range(0, B, 257):
        oa, ov = ch.diff_ik(x0[i], V[i], vmax[i])
        assert abs(alpha[i] - oa) <= 1e-8 * max(1.0, abs(oa))
        if n == 6:
            assert np.abs(v[i] - ov).max() <= 1e-7 * max(1.0, np.abs(ov).max())
    # device path, shared twist and shared vmax
    a2, v2, s2 = r.diff_ik_batch(torch.from_numpy(x0).cuda(), torch.from_numpy(V[0]).cuda(), torch.from_numpy(vmax[0]).cuda())
    a3, v3, s3 = r.diff_ik_batch(x0, np.tile(V[0], (B, 1)), np.tile(vmax[0], (B, 1)))
    assert np.array_equal(a2.cpu().numpy(), a3) and np.array_equal(v2.cpu().numpy(), v3) and np.array_equal(s2.cpu().numpy(), s3)


def test_diff_ik_edge_cases():
    r, ch = robot_and_chain("ur3e")
    # singular configurations (home pose, wrist singularity): the LP is feasible and bounded, the reference returns
    # Some((alpha, v)) (lib.rs:231-238); an unreachable twist gives a zero step, through every entry point
    V = [0.3, 0.1, 0.2, 0.5, 0.4, 0.6]
    import ctypes as C
    lib = ob.load_library()
    arr = lambda a: (C.c_double * len(a))(*a)
    for x0 in (np.zeros(6), np.array([0.3, -1.0, 1.2, 0.4, 0.0, 0.2])):
        alpha, v = r.diff_ik(x0, V, np.ones(6))
        assert alpha == 0.0 and np.all(np.array(v) == 0.0)
        assert ch.diff_ik(x0, V, np.ones(6))[0] == 0.0
        p0 = lib.optik_robot_diff_ik(r._h, arr(x0), arr(V), arr([1.0] * 6))
        assert p0 and np.all(ob._take(p0, 6) == 0.0)
    # a twist inside the range of the singular Jacobian is followed
    from tests_golden_helpers import world_jacobian
    x0 = np.array([0.3, -1.0, 1.2, 0.4, 0.0, 0.2])
    Jw = world_jacobian(ch, x0)
    alpha, v = r.diff_ik(x0, 0.5 * Jw[:, 0], np.ones(6))
    assert alpha == 1.0 and np.abs(Jw @ np.array(v) - 0.5 * Jw[:, 0]).max() < 1e-9
    p = lib.optik_robot_diff_ik(r._h, arr([0.3, -1.0, 1.2, 0.4, 0.7, 0.2]), arr(V), arr([1.0] * 6))
    assert p
    v = ob._take(p, 6)
    assert np.all(np.abs(v) <= 1 + 1e-12)
    # zero twist: alpha = 1, v = 0; empty batch; bad v_max; unsupported size
    a, v0 = r.diff_ik([0.3, -1.0, 1.2, 0.4, 0.7, 0.2], np.zeros(6), np.ones(6))
    assert a == 1.0 and np.all(np.array(v0) == 0)
    a, vv, st = r.diff_ik_batch(np.zeros((0, 6)), np.zeros(6), np.ones(6))
    assert a.shape == (0,) and vv.shape == (0, 6)
    with pytest.raises(ob.OptikError):
        r.diff_ik_batch(np.zeros((1, 6)), np.zeros(6), np.zeros(6))
    snake, _ = robot_and_chain("snake20")
    with pytest.raises(ob.OptikError):
        snake.diff_ik_batch(np.zeros((1, 20)), np.zeros(6), np.ones(20))


@pytest.mark.parametrize("name", ["ur3e", "panda"])
def test_cpp_consumer_end_to_end(name, tmp_path):
    """examples/example.cpp:19-42 protocol through a C++ wrapper that binds the reference's own FFI declarations
    (tests/cpp_wrapper_probe.cpp): default SolverConfig (Speed, max_time 0.1 s, unlimited restarts), random seed ->
    FK-generated target; every returned solution reproduces the target pose inside the limits."""
    import os, re, subprocess
    from conftest import ROOT
    exe = tmp_path / "cpp_wrapper_probe"
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp_wrapper_probe.cpp"),
                           "-o", str(exe), ob.LIB_PATH, "-Wl,-rpath," + os.path.dirname(ob.LIB_PATH)])
    base, ee = ob.ROBOT_LINKS[name]
    p = subprocess.run([str(exe), ob.data_path(name), base, ee, "gpu", "200"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    m = re.search(r"solved=(\d+)/200 accurate=(\d+) diff_ik_ok=(\d+) jac=(\d+)", p.stdout)
    assert m, p.stdout
    solved, accurate, dik, jac = map(int, m.groups())
    assert solved >= 198 and accurate == solved and jac == 6 * (6 if name == "ur3e" else 7)
    assert dik == solved  # diff_ik always has a solution (alpha = 0 at worst)


@pytest.mark.parametrize("name,T,R", [("panda", 20000, 16), ("ur5", 150000, 32), ("ur3e", 12000, 8), ("panda", 10000, 24),
                                      ("panda", 70000, 32)])
def test_dynamic_speed_batches_equal_static_schedule(name, T, R):
    """Speed batches run as dynamic chains (restarts claimed on the device, idle lanes help unsolved targets in
    parallel): q, cost, status, winning restart and the per-target success set are identical to the static
    (target, chunk) schedule and to the lowest-index rule of the twin (lib.rs:409-412 with one thread), on the device
    path and on the host path; evaluations are not compared (helpers run speculative attempts)."""
    import torch
    r, ch = robot_and_chain(name)
    tg, x0, lb, ub = _device_targets(r, ch, T, 11)
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q1, f1, s1, e1 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True, static=True, chunks=1)  # every chain in order
    q2, f2, s2, e2 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True)                          # dynamic chains
    ok = torch.as_tensor(cfg.is_success(s1.cpu().numpy()), device="cuda")
    assert torch.equal(s1, s2) and torch.equal(q1[ok], q2[ok]) and torch.equal(f1[ok], f2[ok])
    assert torch.equal(e1["restart"][ok], e2["restart"][ok])
    assert torch.equal(q1[~ok], q2[~ok])  # unsolved targets: the first attempt's record, in both schedules
    for variant in (1, 2):
        q5, f5, s5 = r.ik_batch(cfg, tg, x0, restarts=R, variant=variant)
        assert torch.equal(s5, s1) and torch.equal(q5[ok], q1[ok])
    q3, f3, s3, e3 = r.ik_batch(cfg, tg.cpu().numpy(), x0.cpu().numpy(), restarts=R, stats=True)  # host path
    q4, f4, s4 = r.ik_batch(cfg, tg.cpu().numpy(), x0.cpu().numpy(), restarts=R, static=True)
    okh = ok.cpu().numpy()
    assert np.array_equal(s3, s1.cpu().numpy()) and np.array_equal(q3[okh], q1[ok].cpu().numpy())
    assert np.array_equal(e3["restart"][okh], e1["restart"][ok].cpu().numpy().astype(np.uint64))
    assert np.array_equal(s4, s3) and np.array_equal(q4[okh], q3[okh])
    # the lowest-index rule against the twin on a sample
    P = O.twin_params(layout=twin_layout(ch))
    tg_h, x0_h = tg.cpu().numpy(), x0.cpu().numpy()
    for t in range(0, T, T // 24):
        ref = O.twin_ik(ch, tg_h[t], x0_h[t], 0, R, "speed", P)
        assert ref["found"] == bool(okh[t])
        if ref["found"]:
            assert int(e3["restart"][t]) == ref["restart"] and np.array_equal(q3[t], ref["q"])


@pytest.mark.parametrize("name,T,R", [("panda", 12000, 6), ("panda", 38000, 32), ("ur3e", 10000, 64)])
def test_dynamic_speed_batches_are_repeatable(name, T, R):
    """Shared targets (record word, restart counters, tickets, in-warp speculation) under repetition: the timing of
    helpers differs from launch to launch, the per-target answer must not -- 12 launches, both column layouts, each
    identical to the static schedule (tools/stress_dynamic.py runs the long version)."""
    import torch
    r, ch = robot_and_chain(name)
    tg, x0, lb, ub = _device_targets(r, ch, T, 29)
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q1, f1, s1, e1 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True, static=True, chunks=1)
    ok = torch.as_tensor(cfg.is_success(s1.cpu().numpy()), device="cuda")
    for rep in range(12):
        q2, f2, s2, e2 = r.ik_batch(cfg, tg, x0, restarts=R, stats=True, variant=1 + rep % 2)
        assert torch.equal(s1, s2) and torch.equal(q1[ok], q2[ok]) and torch.equal(f1[ok], f2[ok]), rep
        assert torch.equal(e1["restart"][ok], e2["restart"][ok]) and torch.equal(q1[~ok], q2[~ok]), rep


@pytest.mark.parametrize("name", ["panda", "ur5"])
def test_evaluator_unaligned_and_ragged_device_inputs(name):
    """The evaluator's TMA tile loads need 16-byte aligned joint vectors and full 32-configuration tiles; a device
    tensor sliced at an odd row (8-byte aligned only) and ragged batch sizes take the plain-load path -- same bits."""
    import torch
    r, ch = robot_and_chain(name)
    B = 4 * 148 * 128 + 77  # more than one tile per warp of the persistent grid, ragged end
    g = torch.Generator(device="cuda").manual_seed(3)
    lb, ub = torch.from_numpy(ch.lb).cuda(), torch.from_numpy(ch.ub).cuda()
    q = (torch.rand((B + 1, ch.n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
    tg = r.eval_batch(q, want=("ee",))["ee"].contiguous()
    full = r.eval_batch(q, tg)
    part = r.eval_batch(q[1:], tg[1:])  # q[1:] starts 8*n bytes in: unaligned for odd n
    assert (q[1:].data_ptr() % 16 != 0) == (ch.n % 2 == 1)
    for k in ("ee", "jac", "f", "grad"):
        assert torch.equal(full[k][1:], part[k]), k
    fk_only = r.eval_batch(q[1:], want=("ee",))["ee"]  # the double-buffered FK-only variant
    assert torch.equal(fk_only, part["ee"])
    for b in (1, 31, 32, 33, 127, 129):
        small = r.eval_batch(q[:b].contiguous(), tg[:b].contiguous())
        for k in ("ee", "jac", "f", "grad"):
            assert torch.equal(small[k], full[k][:b]), (k, b)


# ------------------------------------------------------------------ seeds: known answer + table
def test_chacha8_known_answer_on_device():
    """The kernels' ChaCha8 block function reproduces the published zero-key vector (same check as the oracle's in
    test_oracle_golden.py), and the seed table the solve kernels read equals the oracle's restart seeds bit for bit."""
    from test_oracle_golden import CHACHA8_TC1
    r, ch = robot_and_chain("panda")
    words = r.chacha8_block(np.zeros(8, dtype=np.uint32), 0)
    assert words.astype("<u4").tobytes().hex() == CHACHA8_TC1
    import ctypes as C
    key = (C.c_uint32 * 8)()
    O.lib().oracle_seed_key(C.c_uint64(42), key)
    blk = (C.c_uint32 * 16)()
    for stream in (1, 7, 2 ** 32 + 5):
        O.lib().oracle_chacha8_block(key, 0, stream, blk)
        assert list(r.chacha8_block(np.array(list(key), dtype=np.uint32), stream)) == list(blk)
    for begin, count in ((1, 300), (4090, 20), (70000, 50)):  # inside / across / beyond the resident table
        got = r.restart_seeds(begin, count)
        ref = np.array([ch.restart_seed(i) for i in range(begin, begin + count)])
        assert np.array_equal(got, ref)


# ------------------------------------------------------------------ the solve against an SLSQP-class solver
@pytest.mark.parametrize("name,R", [("panda", 32), ("ur5", 32), ("ur3e", 100)])
def test_solve_against_slsqp_standin(name, R):
    """What the reference runs per restart is NLopt SLSQP (lib.rs:302-356, 372); scipy's SLSQP, configured the same way
    (tests/slsqp_standin.py), stands in for it.  Per seed the two optimisers differ by design;
    what must agree is the RESULT of Robot::ik per target (SURVEY section 7 protocol):
      (1) per-target success agreement >= 99.9 % over 2000 reachable targets, same seeds, <= 32 restarts (100 for
          UR3e, whose per-attempt success is 0.18 for both solvers: with 32 restarts each of them independently misses
          ~0.5 % of the targets);
      (2) both answers satisfy the reference's predicate f < tol_f inside the limits under the golden-pinned oracle;
      (3) polished to f < 1e-20 from the SAME start (the GPU's answer), both solvers land within 1e-6 rad of each other;
      (4) polished from their OWN answers, wherever both sit on the same IK branch they agree to 1e-6 rad; the
          branch-match rate is reported.
    (3) and (4) need isolated solutions: for the 7-DOF arm, whose solutions form a one-parameter family (two polishers
    drift ~1e-4 rad apart along the self-motion direction), they run with the redundancy locked (one joint held)."""
    import slsqp_standin as S
    r, ch = robot_and_chain(name)
    rng = np.random.default_rng(2024)
    T = 2000
    tg = targets_for(ch, rng, T)
    x0 = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R)
    q, f, st = r.ik_batch(cfg, tg, x0, restarts=R)
    ok_gpu = cfg.is_success(st)
    ok_ref = np.zeros(T, dtype=bool)
    q_ref = np.zeros((T, ch.n))
    for t in range(T):
        found, qq, _, _ = S.slsqp_ik(ch, tg[t], x0[t], R)
        ok_ref[t] = found
        if found:
            q_ref[t] = qq
    agree = float((ok_gpu == ok_ref).mean())
    assert agree >= 0.999, (agree, int(ok_gpu.sum()), int(ok_ref.sum()))
    both = np.where(ok_gpu & ok_ref)[0]
    assert len(both) >= 0.995 * T
    for t in both[:: max(1, len(both) // 400)]:  # (2)
        for sol in (q[t], q_ref[t]):
            assert ch.objective(sol, tg[t]) < cfg.tol_f and np.all(sol >= ch.lb) and np.all(sol <= ch.ub)
    if ch.n > 6:
        # joint-space comparisons need isolated solutions: lock the arm's redundancy (joint 3 held at 0.3 rad through
        # lower == upper, in the kernel's clamp and in SLSQP's bounds alike) and solve targets reachable that way
        arr = r.chain()
        arr[2, 12] = arr[2, 13] = 0.3
        r, ch = ob.Robot.from_chain(arr), O.Chain(arr)
        T = 600
        qs = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
        tg = np.stack([ch.fk(qq)[1] for qq in qs])
        x0 = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
        q, f, st = r.ik_batch(cfg, tg, x0, restarts=R)
        ok_l = cfg.is_success(st)
        q_ref = np.zeros((T, ch.n))
        ok_r = np.zeros(T, dtype=bool)
        for t in range(T):
            ok_r[t], qq, _, _ = S.slsqp_ik(ch, tg[t], x0[t], R)
            if ok_r[t]:
                q_ref[t] = qq
        assert (ok_l == ok_r).mean() >= 0.99  # six joints, tight limits: a harder arm for both solvers
        both = np.where(ok_l & ok_r)[0]
    # polish: the GPU path from given starts (one restart = the start itself), SLSQP with scipy
    TOLP = 1e-20
    pcfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=1, tol_f=TOLP)
    sample = both[:600]
    qa, fa, sa = r.ik_batch(pcfg, tg[sample], q[sample], restarts=1, max_evals=64)        # GPU polish of its own answers
    qb, fb, sb = r.ik_batch(pcfg, tg[sample], q_ref[sample], restarts=1, max_evals=64)    # GPU polish of SLSQP's answers
    same_start, own_start, branch = [], [], 0
    for i, t in enumerate(sample):
        if sa[i] != 1:
            continue  # the LM polish stalled above 1e-20 (ill-conditioned pose): not comparable
        okp, qp = S.slsqp_polish(ch, tg[t], q[t], TOLP)
        if okp:
            d = qa[i] - qp
            same_start.append(np.abs(d).max())
        if np.abs(q[t] - q_ref[t]).max() < 0.05:  # same IK branch before polishing
            branch += 1
            oko, qo = S.slsqp_polish(ch, tg[t], q_ref[t], TOLP)
            if oko:
                own_start.append(np.abs(qa[i] - qo).max())
            if sb[i] == 1:
                own_start.append(np.abs(qa[i] - qb[i]).max())
    assert len(same_start) >= 0.6 * len(sample)
    assert max(same_start) <= RAD_TOL, max(same_start)
    assert len(own_start) > 30 and max(own_start) <= RAD_TOL, (len(own_start), max(own_start))
    print(f"\n{name}: per-target success agreement {agree:.4f} (gpu {ok_gpu.mean():.4f}, slsqp {ok_ref.mean():.4f}); "
          f"same-start polish max |dq| {max(same_start):.2e} rad over {len(same_start)}; "
          f"same-branch rate {branch / max(len(sample), 1):.2f}, own-start polish max |dq| "
          f"{max(own_start) if own_start else float('nan'):.2e} rad")


def test_co_resident_passes_on_several_streams_equal_sequential_passes():
    """bench.py's throughput mode: per-attempt passes enqueued on several streams, each on a fraction of the machine
    (opts.blocks), are co-resident on the GPU and share the robot's seed table and scratch pool.  Every pass must give
    exactly the records of the same pass run alone (device path with torch streams, and the asynchronous host path)."""
    import torch
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(21)
    R, NPASS, NSTREAM = 8192, 12, 4
    tg = torch.from_numpy(targets_for(ch, rng, NPASS)).cuda()
    x0 = torch.from_numpy(0.5 * (ch.lb + ch.ub)).cuda()
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
    alone = [tuple(t.clone() for t in r.ik_attempts(cfg, tg[i], x0, R, best=True)) for i in range(NPASS)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(NSTREAM)]
    outs = []
    for i in range(NPASS):
        with torch.cuda.stream(streams[i % NSTREAM]):
            outs.append(r.ik_attempts(cfg, tg[i], x0, R, best=True, blocks=37))
    torch.cuda.synchronize()
    for i in range(NPASS):
        for a, b in zip(alone[i], outs[i]):
            assert torch.equal(a, b), i
    # host path: pinned buffers, OPTIK_BATCH_ASYNC, one library stream per call in flight
    tg_h, x0_h = ob.pinned_empty((NPASS, 8)), ob.pinned_empty(ch.n)
    tg_h[:] = tg.cpu().numpy()
    x0_h[:] = x0.cpu().numpy()
    sets = [((ob.pinned_empty((R, ch.n)), ob.pinned_empty(R), ob.pinned_empty(R, np.int32), ob.pinned_empty(R, np.int32)),
             ob.pinned_empty(ob.RECORD_HEAD + ch.n), ob.Stream(r)) for _ in range(NPASS)]
    for i in range(NPASS):
        r.ik_attempts(cfg, tg_h[i], x0_h, R, best=True, out=sets[i][0], record=sets[i][1], stream=sets[i][2], wait=False, blocks=37)
    for i in range(NPASS):
        sets[i][2].synchronize()
        for a, b in zip(alone[i][:4], sets[i][0]):
            assert np.array_equal(a.cpu().numpy(), b), i
        assert np.array_equal(alone[i][4].cpu().numpy(), sets[i][1]), i


def test_peer_exchange_kernels_single_gpu():
    """optik_gpu_exchange_push / _select (the NVLink best-pick exchange) with four simulated ranks on one GPU: each
    'rank' owns a buffer, every push lands in every buffer, every select returns the record the torch specification of
    the rule (dist.select_candidates) picks; sequence numbers reuse the 32 slots; a missing peer times out as found=-1."""
    import torch
    from optik_b200 import dist as obd
    r, ch = robot_and_chain("panda")
    lib = ob.load_library()
    W, n = 4, ch.n
    L = obd.RECORD_HEAD + n
    nbytes = int(lib.optik_gpu_exchange_bytes(r._h, W))
    assert nbytes == 32 * W * (L * 8 + 8)
    bufs = [torch.zeros(nbytes // 8, dtype=torch.float64, device="cuda") for _ in range(W)]
    peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    rng = np.random.default_rng(3)
    stream = torch.cuda.current_stream().cuda_stream
    for seq in range(1, 70):
        recs = []
        for w in range(W):
            found = float(rng.random() < 0.7)
            rec = obd.pack_candidate(found, float(rng.integers(0, 3)), float(rng.integers(0, 5)), 1e-7,
                                     torch.from_numpy(rng.uniform(-1, 1, n)).cuda())
            recs.append(rec)
            ob._check(lib.optik_gpu_exchange_push(r._h, rec.data_ptr(), peers.data_ptr(), w, W, seq, stream))
        ref = obd.select_candidates(torch.stack(recs), as_tensor=True)
        for w in range(W):
            out = torch.empty(L, dtype=torch.float64, device="cuda")
            ob._check(lib.optik_gpu_exchange_select(r._h, bufs[w].data_ptr(), W, seq, out.data_ptr(), stream))
            assert torch.equal(out, ref), (seq, w)
    # only three of four ranks push sequence 70: the select gives up after its bounded wait
    for w in range(3):
        ob._check(lib.optik_gpu_exchange_push(r._h, recs[w].data_ptr(), peers.data_ptr(), w, W, 70, stream))
    out = torch.empty(L, dtype=torch.float64, device="cuda")
    ob._check(lib.optik_gpu_exchange_select(r._h, bufs[0].data_ptr(), W, 70, out.data_ptr(), stream))
    torch.cuda.synchronize()
    assert out[0].item() == -1.0


def test_ik_fresh_robot_default_config_and_concurrent_calls():
    """(1) The first ik() on a freshly constructed Robot with the reference's DEFAULT config (Speed, max_time = 0.1 s, no
    restart limit, config.rs:52-65) solves a reachable target: the max_time clock must not include the one-time device
    initialisation.  (2) Robot::ik takes &self (lib.rs:241): eight host threads calling ik() on ONE Robot get the same
    answers as serial calls (Speed mode is deterministic: lowest-index converged restart)."""
    import threading
    ch0 = O.Chain(ob.Robot.named("panda").chain())
    rng = np.random.default_rng(77)
    pairs = [(O.pose8_to_matrix(ch0.fk(rng.uniform(ch0.lb, ch0.ub))[1]).tolist(), rng.uniform(ch0.lb, ch0.ub)) for _ in range(64)]
    fresh = ob.Robot.named("panda")
    first = fresh.ik(ob.SolverConfig(), pairs[0][0], pairs[0][1])
    assert first is not None and ch0.objective(np.array(first[0]), O.pose8_from_matrix(np.array(pairs[0][0]))) < 1e-6
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=64)
    serial = [fresh.ik(cfg, m, x0) for m, x0 in pairs]
    assert sum(s is not None for s in serial) >= 63
    out = [None] * len(pairs)

    def work(k):
        for i in range(k, len(pairs), 8):
            out[i] = fresh.ik(cfg, pairs[i][0], pairs[i][1])

    ths = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for a, b in zip(serial, out):
        assert (a is None) == (b is None)
        if a is not None:
            assert a[0] == b[0] and a[1] == b[1]


def test_device_path_flags_clamped_seeds():
    """The reference panics on a seed outside the joint limits (lib.rs:251-254); host-memory calls reject it, device-memory
    calls cannot look without a sync: the kernels clamp the seed and the status word carries
    OPTIK_STATUS_FLAG_SEED_CLAMPED (the success classification ignores the flag)."""
    import torch
    r, ch = robot_and_chain("panda")
    rng = np.random.default_rng(9)
    T = 10000  # enough targets for the dynamic schedule
    tg = torch.from_numpy(targets_for(ch, rng, 64)).cuda().repeat(T // 64 + 1, 1)[:T].contiguous()
    x0 = torch.from_numpy(rng.uniform(ch.lb, ch.ub, size=(T, ch.n))).cuda()
    bad = [3, 77, 9999]
    for t in bad:
        x0[t, t % ch.n] = ch.ub[t % ch.n] + 0.5
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=16)
    for kw in ({}, {"static": True}, {"tile": 8}):
        q, f, st = r.ik_batch(cfg, tg, x0, restarts=16, **kw)
        st = st.cpu().numpy()
        flagged = np.where(st & ob.STATUS_FLAG_SEED_CLAMPED)[0]
        assert list(flagged) == bad, (kw, flagged)
        assert cfg.is_success(st).mean() > 0.99 and np.all((st & ob.STATUS_CODE_MASK) <= 7)
    # per-attempt records: the flag sits on restart 0's record only
    qcfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=64)
    for tile in (0, 8):
        qa, fa, sa, ea, rec = r.ik_attempts(qcfg, tg[3], x0[3], 64, tile=tile, best=True)
        sa = sa.cpu().numpy()
        assert sa[0] & ob.STATUS_FLAG_SEED_CLAMPED and not np.any(sa[1:] & ob.STATUS_FLAG_SEED_CLAMPED)
        assert float(rec[4]) == float(int(rec[4]) & ob.STATUS_CODE_MASK)
        qb, fb, sb, eb = r.ik_attempts(qcfg, tg[5], x0[5], 64, tile=tile)
        assert not np.any(sb.cpu().numpy() & ob.STATUS_FLAG_SEED_CLAMPED)


@pytest.mark.parametrize("T", [600, 12000])
def test_batch_default_budget_runs_until_max_time(T):
    """The reference's default budget -- no restart limit, max_time bounds the call (config.rs:52-65, lib.rs:260-277) --
    on the batched path: reachable targets are solved, unreachable ones keep drawing restarts until the deadline
    (tests/test_ik.rs:24-43 at batch size), through the static (small T) and the dynamic (large T) schedule."""
    import time
    r, ch = robot_and_chain("ur3e")
    rng = np.random.default_rng(21)
    tg = targets_for(ch, rng, 200)
    tg = np.tile(tg, (T // 200, 1))
    far = np.arange(0, T, 50)
    tg[far, 4:7] = [100.0, 100.0, 100.0]  # unreachable translations
    x0 = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
    cfg = ob.SolverConfig(max_time=0.05)  # Speed, max_restarts = u64::MAX
    r.ik_batch(ob.SolverConfig(max_time=0.0, max_restarts=2), tg[:8], x0[:8])  # warm the context
    t0 = time.perf_counter()
    q, f, st = r.ik_batch(cfg, tg, x0)
    dt = time.perf_counter() - t0
    ok = cfg.is_success(st)
    assert not ok[far].any() and 0.04 <= dt <= 0.25, dt
    near = np.setdiff1d(np.arange(T), far)
    assert ok[near].mean() > 0.999
    for t in near[:: max(1, len(near) // 64)]:
        assert ch.objective(q[t], tg[t]) < cfg.tol_f
    with pytest.raises(ob.OptikError, match="forever"):
        r.ik_batch(ob.SolverConfig(max_time=0.0, max_restarts=0) if False else _no_budget(), tg[:4], x0[:4])


def _no_budget():
    c = ob.SolverConfig(max_time=0.1)
    c.max_time = 0.0  # bypass the constructor's check (crates/optik-py/src/lib.rs:45-47): the C ABI refuses it too
    return c
