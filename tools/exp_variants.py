"""Kernel-variant experiments: build liboptik_b200 under several -D switches (here, on the CPU box), then time each
variant on the GPU box in its own process.

    python tools/exp_variants.py build  name=DEF1,DEF2 ...     # -> optik_b200/lib/exp_<name>.so (ships with gpurun)
    python tools/exp_variants.py run [what]                     # on the GPU box: every exp_*.so + the product library
"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBDIR = os.path.join(ROOT, "optik_b200", "lib")

if sys.argv[1] == "build":
    import optik_b200.build as B
    from concurrent.futures import ThreadPoolExecutor
    jobs = []
    for spec in sys.argv[2:]:
        name, _, defs = spec.partition("=")
        jobs.append((name, [d for d in defs.split(",") if d]))
    with ThreadPoolExecutor(4) as ex:
        for name, out in zip([j[0] for j in jobs], ex.map(lambda j: B.build(force=True, defines=j[1], out=os.path.join(LIBDIR, f"exp_{j[0]}.so")), jobs)):
            print(name, out)
elif sys.argv[1] == "run":
    what = sys.argv[2] if len(sys.argv) > 2 else "step"
    libs = [os.path.join(LIBDIR, "liboptik_b200.so")] + sorted(glob.glob(os.path.join(LIBDIR, "exp_*.so")))
    for lib in libs:
        print("=====", os.path.basename(lib), flush=True)
        subprocess.call([sys.executable, os.path.abspath(__file__), "one", lib, what])
elif sys.argv[1] == "one":
    import optik_b200 as ob
    ob.LIB_PATH = sys.argv[2]
    sys.argv = [sys.argv[0], sys.argv[3]]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import exp_r2
    import torch
    what = sys.argv[1]
    dev = exp_r2.dev
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # fingerprint of the results (variants must not change a bit)
    import hashlib, numpy as np
    r = ob.Robot.named("panda")
    lb, ub = map(np.array, r.joint_limits())
    tg = r.eval_batch(torch.from_numpy(np.random.default_rng(7).uniform(lb, ub, size=(1, 7))).to(dev), want=("ee",))["ee"][0].contiguous()
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=8192)
    h = hashlib.sha1()
    for v in (1, 2):
        out = r.ik_attempts(cfg, tg, torch.from_numpy(0.5 * (lb + ub)).to(dev), 8192, variant=v)
        for a in out[:4]:
            h.update(a.cpu().numpy().tobytes())
    print("fingerprint", h.hexdigest(), flush=True)
    if what in ("step", "all"):
        for v in (2, 1):
            exp_r2.step_times(65536, v, flush=flush, reps=20)
        exp_r2.step_times(1 << 20, 1, reps=3)
        exp_r2.step_times(1 << 20, 2, reps=3)
    if what in ("batch", "all"):
        for T in (1 << 14, 1 << 16, 1 << 18, 1 << 20):
            exp_r2.batch("panda", T, 32)
        exp_r2.batch("ur5", 1 << 20, 32)
        exp_r2.batch("panda", 1 << 18, 32, mode="quality")
