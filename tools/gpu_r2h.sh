set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
import time, ctypes as C, numpy as np, threading
import optik_b200 as ob
lib = ob.load_library()
r = ob.Robot.named("panda")
lb, ub = map(np.array, r.joint_limits())
rng = np.random.default_rng(42)
cfg = ob.SolverConfig()._c()
pairs = []
for _ in range(2020):
    qs, x0 = rng.uniform(lb, ub), rng.uniform(lb, ub)
    pairs.append((np.array(r.fk(qs)).T.copy(), x0.copy()))
dp = C.POINTER(C.c_double)
def run(pp):
    ok = 0
    for m, x0 in pp:
        p = lib.optik_robot_ik(r._h, C.byref(cfg), m.ctypes.data_as(dp), x0.ctypes.data_as(dp))
        if p: lib.free(p); ok += 1
    return ok
run(pairs[:20])
t0 = time.perf_counter(); ok = run(pairs[20:]); dt = time.perf_counter() - t0
print(f"single thread: {dt/2000*1e6:.1f} us/call ok {ok}/2000")
for nt in (2, 4, 8):
    ths = [threading.Thread(target=run, args=(pairs[20 + k::nt],)) for k in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
    print(f"{nt} threads: {2000/dt:.0f} calls/s ({dt/2000*1e6:.1f} us/call aggregate)")
PY
