"""Single-call latency (BASELINE config 1 protocol: default SolverConfig through optik_robot_ik) for the library in argv[1]."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import optik_b200 as ob
ob.LIB_PATH = sys.argv[1]
r = ob.Robot.named("panda")
lb, ub = map(np.array, r.joint_limits())
rng = np.random.default_rng(42)
N = 3000
pairs = [(np.array(r.fk(rng.uniform(lb, ub))).tolist(), list(rng.uniform(lb, ub))) for _ in range(N)]
cfg = ob.SolverConfig()
for m, x0 in pairs[:50]:
    r.ik(cfg, m, x0)
t0 = time.perf_counter(); ok = 0
for m, x0 in pairs:
    ok += r.ik(cfg, m, x0) is not None
dt = time.perf_counter() - t0
print(f"{sys.argv[1].split('/')[-1]}: {dt / N * 1e6:.2f} us/call (python wrapper included), success {ok / N:.4f}", flush=True)
