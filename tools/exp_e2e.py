"""e2e pipelining experiment (one GPU): BASELINE config 2 passes through Robot.ik_attempts with pinned host buffers and
D calls in flight, for the library given by OPTIK_EXP_LIB (default: the product library).
    python tools/exp_e2e.py [depths...]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optik_b200 as ob
import optik_b200.dist as obd
if os.environ.get("OPTIK_EXP_LIB"):
    ob.LIB_PATH = os.environ["OPTIK_EXP_LIB"]
dev = torch.device("cuda", 0)
robot = ob.Robot.named("panda")
n = 7
lb, ub = map(np.array, robot.joint_limits())
R = 65536
Ke = 1500
rng = np.random.default_rng(42)
qstar = torch.from_numpy(rng.uniform(lb, ub, size=(Ke, n))).to(dev)
targets = robot.eval_batch(qstar, want=("ee",))["ee"].contiguous()
cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R)
tg_host = ob.pinned_empty((Ke, 8)); tg_host[:] = targets.cpu().numpy()
x0_host = ob.pinned_empty(n); x0_host[:] = 0.5 * (lb + ub)
BLOCKS = int(os.environ.get("OPTIK_EXP_BLOCKS", "0"))
for D in [int(a) for a in sys.argv[1:]] or [2, 3, 4]:
    sets = [((ob.pinned_empty((R, n)), ob.pinned_empty(R), ob.pinned_empty(R, np.int32), ob.pinned_empty(R, np.int32)),
             ob.pinned_empty(obd.RECORD_HEAD + n), ob.Stream(robot)) for _ in range(D)]
    def submit(s, k):
        hrec, hrecord, stream = sets[k]
        robot.ik_attempts(cfg, tg_host[s], x0_host, R, best=True, out=hrec, record=hrecord, stream=stream, wait=False, blocks=BLOCKS)
    def done(k):
        sets[k][2].synchronize()
        return int((sets[k][0][2] == 1).sum())
    for k in range(D): submit(0, k)
    for k in range(D): done(k)
    conv = 0
    t0 = time.perf_counter()
    for s in range(Ke):
        k = s % D
        if s >= D: conv += done(k)
        submit(s, k)
    for s in range(max(Ke - D, 0), Ke): conv += done(s % D)
    dt = time.perf_counter() - t0
    print(f"{os.path.basename(ob.LIB_PATH)} blocks {BLOCKS} depth {D}: {dt / Ke * 1e3:.4f} ms/pass  e2e {conv / dt:.4e} converged attempts/s", flush=True)
