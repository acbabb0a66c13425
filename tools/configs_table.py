"""Measures the five BASELINE.json configs on one GPU (config 5 = its per-GPU shard) and prints a markdown table.
Device-side timing (CUDA events), inputs resident in HBM; success = status code + re-evaluation by the evaluator kernel."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob

dev = torch.device("cuda", 0)
rows = []

def targets(robot, T, seed):
    lb, ub = [torch.tensor(x, dtype=torch.float64, device=dev) for x in robot.joint_limits()]
    g = torch.Generator(device=dev).manual_seed(seed)
    q = torch.rand((T, robot.num_positions()), dtype=torch.float64, device=dev, generator=g) * (ub - lb) + lb
    x0 = torch.rand((T, robot.num_positions()), dtype=torch.float64, device=dev, generator=g) * (ub - lb) + lb
    return robot.eval_batch(q, want=("ee",))["ee"].contiguous(), x0.contiguous(), lb, ub

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out

def gate(robot, cfg, q, tg, st, lb, ub):
    fe = robot.eval_batch(q, tg, want=("f",))["f"]
    ok = torch.as_tensor(cfg.is_success(st.cpu().numpy()), device=dev) & (fe < cfg.tol_f) & ((q >= lb) & (q <= ub)).all(dim=1)
    return float(ok.double().mean())

# config 1: Panda single solve through the C ABI, Speed, default config (plumbing / latency)
r = ob.Robot.named("panda")
tg, x0, lb, ub = targets(r, 2000, 1)
tgh, x0h = tg.cpu().numpy(), x0.cpu().numpy()
cfg = ob.SolverConfig()  # Speed, max_time 0.1, tol_f 1e-6
lib = ob.load_library()
mats = [np.ascontiguousarray(np.array(ob._rows_from_pose8(p)).T).ravel() for p in tgh]  # column-major 4x4
cc = cfg._c()
r.ik(cfg, ob._rows_from_pose8(tgh[0]), x0h[0])  # warm-up
t0 = time.perf_counter(); ok1 = 0
for i in range(2000):
    p = lib.optik_robot_ik(r._h, C.byref(cc), mats[i].ctypes.data_as(C.POINTER(C.c_double)), x0h[i].ctypes.data_as(C.POINTER(C.c_double)))
    if p: ok1 += 1; lib.free(p)
dt = time.perf_counter() - t0
rows.append(("1", "Panda single ik() via optik_robot_ik (C ABI), Speed, defaults", f"{dt/2000*1e6:.0f} us/call", f"{2000/dt:.3e} calls/s", f"{ok1/2000:.4f}", "host wall clock, 2000 sequential calls"))

# config 2: Panda, 65536 seeds to one target
cfgq = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=65536)
x0m = (0.5 * (lb + ub)).contiguous()
for tile in (1, 8, 32):
    ms, out = timed(lambda: r.ik_attempts(cfgq, tg[0], x0m, 65536, tile=tile, best=True))
    conv = float((out[2] == 1).double().mean())
    rows.append(("2", f"Panda, 65536 seeds -> 1 target, Quality, lanes/seed={tile}", f"{ms:.3f} ms", f"{65536*conv/ms*1e3:.3e} converged/s ({65536/ms*1e3:.3e} attempts/s)", f"{conv:.4f} per attempt", "per-restart records + selection"))

# config 3: UR5, 1 Mi targets, Speed, <= 32 restarts
u = ob.Robot.named("ur5")
tg3, x03, lb3, ub3 = targets(u, 1 << 20, 3)
cfg3 = ob.SolverConfig(max_time=0.0, max_restarts=32)
ms, out = timed(lambda: u.ik_batch(cfg3, tg3, x03, restarts=32))
s3 = gate(u, cfg3, out[0], tg3, out[2], lb3, ub3)
rows.append(("3", "UR5, 1 Mi independent targets, Speed, <=32 restarts", f"{ms:.2f} ms", f"{(1<<20)*s3/ms*1e3:.3e} solves/s", f"{s3:.5f} per target", "lowest-index converged restart per target"))

# config 4: snake20, 262144 seeds
sn = ob.Robot.named("snake20")
tg4, x04, lb4, ub4 = targets(sn, 4, 4)
cfg4 = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=262144)
ms, out = timed(lambda: sn.ik_attempts(cfg4, tg4[0], (0.5 * (lb4 + ub4)).contiguous(), 262144, best=True))
conv = float((out[2] == 1).double().mean())
rows.append(("4", "20-DOF snake, 262144 seeds -> 1 target, Quality (tile kernel, 32 lanes)", f"{ms:.3f} ms", f"{262144*conv/ms*1e3:.3e} converged/s", f"{conv:.4f} per attempt", "6x6 dual solve: cost independent of n"))

# config 5: Panda Quality 256 restarts x 1 Mi targets over 8 GPUs -> one GPU's shard = 131072 targets
tg5, x05, _, _ = targets(r, 131072, 5)
cfg5 = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=256)
ms, out = timed(lambda: r.ik_batch(cfg5, tg5, x05, restarts=256), reps=2)
s5 = gate(r, cfg5, out[0], tg5, out[2], lb, ub)
rows.append(("5 (1/8 shard)", "Panda Quality, 256 restarts x 131072 targets (one GPU's shard of 1 Mi)", f"{ms:.1f} ms", f"{131072*s5/ms*1e3:.3e} solves/s ({131072*256/ms*1e3:.3e} attempts/s)", f"{s5:.5f} per target", "arg-min ||q-x0|| over converged restarts"))

print("| config | workload | time | throughput | success | note |")
print("|---|---|---|---|---|---|")
for row in rows:
    print("| " + " | ".join(row) + " |")
