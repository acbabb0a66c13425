#!/usr/bin/env python
"""bench.py -- headline benchmark of the IK hot path (BASELINE.json metric) + the other BASELINE configs.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, all host threads)

Headline workload (BASELINE configs[1]; one "pass"): Panda 7-DOF, 65 536 random-restart seeds to ONE reachable target per
GPU, SolutionMode::Quality (every restart runs, lib.rs:398-407), tol_f = 1e-6, max_time = 0.  Rank g runs the restart
range [g*65536, (g+1)*65536) (weak scaling); each rank selects its best candidate on the device and ONE NCCL all-gather
of an 15-double record per rank picks the global best.  A fresh target is used every pass.  A timed STEP is
`passes_per_step` passes, sized so that the K timed steps cover >= 1 s of device time (clock samples need it).

`value`: success-gated converged restart attempts per second (each converged attempt is RE-EVALUATED by the evaluator
kernel outside the timed region and must satisfy the reference's predicate f(q) < tol_f with lb <= q <= ub; a sample is
re-checked by the golden-pinned CPU oracle), device time (CUDA events, max over ranks).  The per-TARGET metric of
SURVEY section 8(d) -- IK problems solved per second -- is the `per_target` block (Panda, 1 Mi independent reachable
targets, Speed, <= 32 restarts, through optik_gpu_ik_batch) and `configs` carries BASELINE configs 1, 3, 4 (and 5 on
N > 1 GPUs), each next to the CPU port on the same workload (N = 1 only).

`value` times every pass ALONE on the whole GPU.  `e2e` is the same pass through the host-buffer API in throughput mode
(asynchronous calls, several in flight, each on a fraction of the machine: E2E_* below) and `device_throughput_mode` its
device-only twin.  `roofline` = the evaluator kernel against the measured HBM copy peak, with the measured ceiling of a
pure streaming kernel at the same read/write mix (`mix_ceiling`) and the oracle's CPU evaluator (`cpu_baseline`) beside
it; `roofline_solve` = the solve kernel against an fp64 FMA peak measured in the same run.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEEDS_PER_GPU = 65536
MIN_TIMED_SECONDS = 1.0
CPU_BASELINE_PASSES = 32  # cpu_baseline sample: 32 full passes (2 Mi attempts, ~20 CPU-seconds on 16 cores)
E2E_DEPTH_SINGLE = int(os.environ.get("OPTIK_BENCH_E2E_DEPTH", "10"))  # host-buffer calls in flight on one GPU
# e2e throughput mode: every pass is launched on a FRACTION of the machine (opts.blocks = SMs / E2E_SM_DIV: a third of the
# SM count = a sixth of the resident block slots) so that several passes are co-resident and one pass's straggler tail runs
# under the next passes' bulk (measured on B200, tools/exp_e2e.py: full-machine launches 0.208 ms/pass at any depth, 74
# blocks with 8 passes in flight 0.150, 49 blocks with 10 in flight 0.143, 29 blocks with 16 in flight 0.137 ms/pass; the
# device-timed `value` stays the isolated full-machine pass).  Several GPUs: half of the SM count, 8 passes in flight.
E2E_SM_DIV = int(os.environ.get("OPTIK_BENCH_E2E_SM_DIV", "3"))
E2E_SM_DIV_MULTI = 2
PT_DEPTH = 3  # per_target e2e: host-buffer batch calls in flight
E2E_DEPTH_MULTI = 8  # host-buffer steps in flight per rank when a collective sits inside the step (N > 1)
ROBOT = "panda"
TOL_F = 1e-6
METRIC = "IK solves/sec (success-gated, Panda 7-DOF)"
UNIT = "solves/s (1 solve = 1 converged restart attempt; IK problems/s: see per_target)"
LINKS = {"panda": ("panda_link0", "panda_link8"), "ur5": ("base_link", "ee_link"), "ur3e": ("ur_base_link", "ur_ee_link"),
         "snake20": ("seg0", "tip")}
PER_TARGET_T = 1 << 20
PER_TARGET_R = 32
CPU_FLAGS = "gcc -O3 -mfma -ffp-contract=off (bit-exact twin arithmetic: own sin/cos/atan), no -march=native"


def workload_config(n_gpus):
    """The workload, identical for both arms (what differs per arm is in `impl_detail`)."""
    return {
        "workload": "configs[1]: Panda 7-DOF, 65536 random-restart seeds to one target per GPU per pass, "
                    "SolutionMode::Quality, tol_f=1e-6, max_time=0, fresh reachable target every pass",
        "robot": ROBOT, "dof": 7, "seeds_per_gpu_per_pass": SEEDS_PER_GPU, "targets_per_pass": 1,
        "solve_definition": "one restart attempt that converged (f<tol_f inside the joint limits, re-verified); the "
                            "per-target metric (IK problems solved per second) is the per_target block",
        "parallelism": f"restart-range sharding over {n_gpus} GPU(s); one exchange of 15 doubles/rank/pass "
                       "for the Quality best-pick" if n_gpus > 1 else "single GPU",
    }


def impl_detail(tile, passes_per_step):
    return {
        "passes_per_step": passes_per_step,
        "lanes_per_seed": tile, "layout": ("thread-per-seed kernel (solve_t1_kernel)" if tile == 1 else
                                           f"tile kernel, {tile} lanes per seed (solve_kernel<{tile}>)" if tile else
                                           "cpu port: pthread workers over the fp64 LM twin"),
        "l2": "flushed (256 MiB write) before every timed pass; passes timed individually with CUDA events" if tile else "n/a",
    }


def oracle_chain(name):
    """The checker's own URDF loader (oracle/urdf_chain.py): the reference arm never maps the product library."""
    from oracle import oracle as O
    base, ee = LINKS[name]
    return O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", name + ".urdf")).read(), base, ee)


def seeded_q(lb, ub, count, seed=42):
    """Seeded joint vectors uniform in the limits -> reachable targets (examples/example.rs:24-26 protocol)."""
    return np.random.default_rng(seed).uniform(lb, ub, size=(count, len(lb)))


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max, self.error, self.power = [], set(), None, None, []
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # pragma: no cover
            self.error = repr(e)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        d = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
             "reasons": sorted(self.reasons), "samples": len(self.samples),
             "power_w_max": float(max(self.power)) if self.power else None}
        if self.error:
            d["error"] = self.error
        return d


# ----------------------------------------------------------------------------------------------- CPU port (oracle)
def cpu_reference_passes(passes, warmup, threads, seeds=SEEDS_PER_GPU):
    """The reference's CPU path for the headline workload: the oracle port of the rayon restart loop (oracle/ref_loop.c
    over the fp64 solver twin) with `threads` host threads.  Returns (converged, seconds, attempts)."""
    from oracle import oracle as O
    ch = oracle_chain(ROBOT)
    qstar = seeded_q(ch.lb, ch.ub, passes + warmup)
    x0 = 0.5 * (ch.lb + ch.ub)
    P = O.twin_params(layout=1)
    conv = att = 0
    total = 0.0
    for s in range(passes + warmup):
        tgt = ch.fk(qstar[s])[1]
        t0 = time.perf_counter()
        res = O.ref_ik_threaded(ch, tgt, x0, 0, seeds, "quality", threads, params=P)
        dt = time.perf_counter() - t0
        if s >= warmup:
            conv += res["converged"]
            att += res["attempts"]
            total += dt
    return conv, total, att


def cpu_per_target(name, T, R, threads, seed=7):
    """CPU port on the per-target workload: T independent reachable targets, Speed, <= R restarts, one target per
    worker thread (Robot::ik with set_parallelism(1) per target: no speculative attempt is wasted).  -> dict."""
    from oracle import oracle as O
    ch = oracle_chain(name)
    rng = np.random.default_rng(seed)
    tg = np.stack([ch.fk(q)[1] for q in rng.uniform(ch.lb, ch.ub, size=(T, ch.n))])
    x0 = rng.uniform(ch.lb, ch.ub, size=(T, ch.n))
    P = O.twin_params(layout=1 if ch.n <= 8 else 0)
    O.ref_batch_threaded(ch, tg[:64], x0[:64], R, "speed", threads, params=P)
    t0 = time.perf_counter()
    q, f, found = O.ref_batch_threaded(ch, tg, x0, R, "speed", threads, params=P)
    dt = time.perf_counter() - t0
    return {"value": float(found.sum() / dt), "unit": "targets solved/s", "cores": threads, "kind": "port",
            "success_rate": float(found.mean()),
            "sample": f"{T} {name} targets, Speed, <= {R} restarts, one target per worker thread ({dt:.2f} s wall); {CPU_FLAGS}"}


def cpu_single_calls(name, calls, seed=42):
    """BASELINE config 1 on the CPU port: one ik() per call, ONE thread, Speed, restarts until success (cap 1000)."""
    from oracle import oracle as O
    ch = oracle_chain(name)
    rng = np.random.default_rng(seed)
    P = O.twin_params(layout=1)
    pairs = [(ch.fk(rng.uniform(ch.lb, ch.ub))[1], rng.uniform(ch.lb, ch.ub)) for _ in range(calls)]
    ok = 0
    t0 = time.perf_counter()
    for tgt, x0 in pairs:
        ok += O.twin_ik(ch, tgt, x0, 0, 1000, "speed", P)["found"]
    dt = time.perf_counter() - t0
    return {"us_per_call": dt / calls * 1e6, "success_rate": ok / calls, "cores": 1, "kind": "port",
            "sample": f"{calls} calls through ctypes (twin_ik, one thread); {CPU_FLAGS}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = args.steps
    conv, secs, att = cpu_reference_passes(steps, args.warmup, threads)
    value = conv / secs
    sample = (f"{steps} passes x {SEEDS_PER_GPU} seeds to one target each (the full per-GPU pass), oracle port of the rayon "
              f"restart loop (lib.rs:297-413) over the fp64 LM twin on all {threads} logical cores (the reference's default, "
              f"lib.rs:45; its README advises cores/2); {CPU_FLAGS}; the Rust reference (NLopt SLSQP) cannot be built in this "
              "image -- scipy's SLSQP needs ~3x the objective evaluations per converged attempt of this port "
              "(tests/experiments/exp_slsqp_calibration.py), so the port is the faster baseline")
    pt = None
    if not args.no_cpu_baseline:
        pt = cpu_per_target(ROBOT, 1 << 17, PER_TARGET_R, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus), "impl_detail": impl_detail(0, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "attempts_per_s": att / secs, "gpu_launches": 0, "per_target": pt,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- product arm helpers
def device_targets(robot, name, T, seed, dev):
    """T reachable targets + uniform seeds, generated on the device (outside every timed region)."""
    import torch
    n = robot.num_positions()
    lb, ub = [torch.tensor(x, dtype=torch.float64, device=dev) for x in robot.joint_limits()]
    g = torch.Generator(device=dev).manual_seed(seed)
    qs = torch.rand((T, n), dtype=torch.float64, device=dev, generator=g) * (ub - lb) + lb
    x0 = (torch.rand((T, n), dtype=torch.float64, device=dev, generator=g) * (ub - lb) + lb).contiguous()
    tg = robot.eval_batch(qs, want=("ee",))["ee"].contiguous()
    return tg, x0, lb, ub


def gate_targets(robot, cfg, q, st, tg, lb, ub):
    """SURVEY 8(d) success gate per target: the returned q, re-evaluated, has f < tol_f inside the limits."""
    import torch
    fe = robot.eval_batch(q, tg, want=("f",))["f"]
    ok = torch.as_tensor(cfg.is_success(st.cpu().numpy()), device=q.device)
    return int((ok & (fe < cfg.tol_f) & ((q >= lb) & (q <= ub)).all(dim=1)).sum()), int(ok.sum())


def timed_batch(robot, cfg, tg, x0, lb, ub, restarts, flush, reps=3, **kw):
    """Device-timed Robot.ik_batch over resident inputs: best of `reps` after one warm-up.  -> (ms, verified, claimed)."""
    import torch
    best, q, st = 1e30, None, None
    for i in range(reps + 1):
        flush.fill_(i & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        q, f, st = robot.ik_batch(cfg, tg, x0, restarts=restarts, **kw)
        b.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, a.elapsed_time(b))
    ver, claimed = gate_targets(robot, cfg, q, st, tg, lb, ub)
    return best, ver, claimed


def per_target_block(ob, robot, dev, flush, rank, world, cpu):
    """Panda, 1 Mi independent reachable targets per GPU, Speed, <= 32 restarts, through optik_gpu_ik_batch."""
    import torch
    T, R = PER_TARGET_T, PER_TARGET_R
    cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=R, tol_f=TOL_F)
    tg, x0, lb, ub = device_targets(robot, ROBOT, T, 1000 + rank, dev)
    ms, ver, claimed = timed_batch(robot, cfg, tg, x0, lb, ub, R, flush)
    n = robot.num_positions()
    # e2e: pinned host buffers, H2D of targets + seeds and D2H of q / cost / status inside the timed region, PT_DEPTH calls
    # in flight, each on its own library stream (OPTIK_BATCH_ASYNC)
    tg_h, x0_h = ob.pinned_empty((T, 8)), ob.pinned_empty((T, n))
    tg_h[:] = tg.cpu().numpy()
    x0_h[:] = x0.cpu().numpy()
    sets = [((ob.pinned_empty((T, n)), ob.pinned_empty(T), ob.pinned_empty(T, np.int32)), ob.Stream(robot)) for _ in range(PT_DEPTH)]
    steps = 24
    solved = 0

    def finish(k):
        out, stream = sets[k]
        stream.synchronize()
        return int(cfg.is_success(out[2]).sum())

    for k in range(PT_DEPTH):  # warm the streams / pools
        robot.ik_batch(cfg, tg_h, x0_h, restarts=R, out=sets[k][0], stream=sets[k][1], wait=False)
    for k in range(PT_DEPTH):
        finish(k)
    t0 = time.perf_counter()
    for s in range(steps):
        k = s % PT_DEPTH
        if s >= PT_DEPTH:
            solved += finish(k)
        robot.ik_batch(cfg, tg_h, x0_h, restarts=R, out=sets[k][0], stream=sets[k][1], wait=False)
    for s in range(steps - PT_DEPTH, steps):
        solved += finish(s % PT_DEPTH)
    e2e_s = time.perf_counter() - t0
    blk = {
        "workload": f"Panda 7-DOF, {T} independent reachable targets per GPU (uniform seeds), SolutionMode::Speed, <= {R} "
                    "restarts, tol_f=1e-6, one optik_gpu_ik_batch call (dynamic chains, one launch)",
        "value": ver / (ms * 1e-3), "unit": "targets solved/s (returned q re-verified: f<tol_f inside the limits)",
        "ms_per_call": ms, "targets": T, "success_rate": ver / T, "verified_equals_claimed": ver == claimed,
        "e2e": {"value": solved / e2e_s, "unit": "targets solved/s", "h2d_bytes_per_step": T * (64 + 8 * n),
                "d2h_bytes_per_step": T * (8 * n + 8 + 4), "steps": steps, "ms_per_step": e2e_s / steps * 1e3,
                "api": "Robot.ik_batch(pinned host buffers, stream=, wait=False) -> optik_gpu_ik_batch + OPTIK_BATCH_ASYNC, "
                       "two calls in flight"},
        "cpu_baseline": cpu,
    }
    del tg, x0, tg_h, x0_h, sets
    return blk


def other_configs(ob, robot, dev, flush, rank, world, with_cpu):
    """BASELINE configs 1, 3, 4 on this rank's GPU (5 separately: it needs the process group)."""
    import torch
    import ctypes as C
    out = {}
    # ---- config 1: single ik() calls through the reference's own C symbol, default SolverConfig
    lib = ob.load_library()
    lbn, ubn = map(np.array, robot.joint_limits())
    rng = np.random.default_rng(42)
    calls = 2000
    cfg1 = ob.SolverConfig()._c()  # Speed, max_time 0.1 s, no restart limit (config.rs:52-65)
    pairs = []
    for _ in range(calls + 20):
        qs, x0 = rng.uniform(lbn, ubn), rng.uniform(lbn, ubn)
        m = np.array(robot.fk(qs)).T.copy()  # column-major 4x4 for optik_robot_ik
        pairs.append((m, x0.copy()))
    dp = C.POINTER(C.c_double)
    ok = 0
    for i, (m, x0) in enumerate(pairs):
        if i == 20:
            t0 = time.perf_counter()
        p = lib.optik_robot_ik(robot._h, C.byref(cfg1), m.ctypes.data_as(dp), x0.ctypes.data_as(dp))
        if p:
            lib.free(p)
            ok += i >= 20
    dt = time.perf_counter() - t0
    out["config1"] = {"workload": "configs[0]: Panda single IK solve per call through optik_robot_ik (the reference's C symbol), "
                                  "default SolverConfig (Speed, max_time 0.1 s, unlimited restarts), examples/example.rs protocol",
                      "us_per_call": dt / calls * 1e6, "success_rate": ok / calls, "calls": calls,
                      "cpu_baseline": cpu_single_calls(ROBOT, 2000) if with_cpu else None}
    # ---- config 3: UR5, 1 Mi targets, Speed
    ur5 = ob.Robot.named("ur5")
    ur5.set_device(dev.index)
    cfg3 = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=32, tol_f=TOL_F)
    tg, x0, lb, ub = device_targets(ur5, "ur5", 1 << 20, 300 + rank, dev)
    ms, ver, claimed = timed_batch(ur5, cfg3, tg, x0, lb, ub, 32, flush)
    out["config3"] = {"workload": "configs[2]: UR5 6-DOF, 1 Mi independent reachable targets, Speed, <= 32 restarts",
                      "value": ver / (ms * 1e-3), "unit": "targets solved/s", "ms_per_call": ms, "success_rate": ver / (1 << 20),
                      "cpu_baseline": cpu_per_target("ur5", 1 << 15, 32, os.cpu_count() or 1) if with_cpu else None}
    del tg, x0
    # ---- config 4: 20-DOF snake, 262 144 seeds to one target (tile kernel, one warp per seed)
    snake = ob.Robot.named("snake20")
    snake.set_device(dev.index)
    lbs, ubs = map(np.array, snake.joint_limits())
    R4 = 262144
    cfg4 = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R4, tol_f=TOL_F)
    tgs = snake.eval_batch(torch.from_numpy(seeded_q(lbs, ubs, 4)).to(dev), want=("ee",))["ee"].contiguous()
    x0s = torch.from_numpy(0.5 * (lbs + ubs)).to(dev)
    best, conv = 1e30, 0
    for i in range(4):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        q, f, st, ev = snake.ik_attempts(cfg4, tgs[i], x0s, R4)
        b.record()
        torch.cuda.synchronize()
        if i and a.elapsed_time(b) < best:
            best = a.elapsed_time(b)
            fe = snake.eval_batch(q, tgs[i], want=("f",))["f"]
            conv = int(((st == 1) & (fe < TOL_F)).sum())
    out["config4"] = {"workload": "configs[3]: 20-DOF snake, 262144 random-restart seeds to one target, Quality (tile kernel)",
                      "value": conv / (best * 1e-3), "unit": "converged attempts/s (re-verified)", "ms_per_call": best,
                      "success_rate_per_attempt": conv / R4}
    return out


def config5_block(ob, robot, dev, flush, rank, world):
    """BASELINE config 5: Panda Quality, 256 restarts per target, targets sharded over the ranks (131 072 per GPU = 1 Mi
    on 8 GPUs), ONE all-gather of the per-target results."""
    import torch
    import torch.distributed as dist
    from optik_b200 import dist as obd
    T, R = 131072, 256
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R, tol_f=TOL_F)
    tg, x0, lb, ub = device_targets(robot, ROBOT, T, 500 + rank, dev)
    best = 1e30
    for i in range(3):
        flush.fill_(i)
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        q, f, st = obd.ik_batch_target_sharded(robot, cfg, tg, x0, R, rank=rank, world=world)
        b.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, a.elapsed_time(b))
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nq = robot.num_positions()
    mine = slice(rank * T, (rank + 1) * T)
    ver, claimed = gate_targets(robot, cfg, q[mine].contiguous(), st[mine].contiguous(), tg, lb, ub)
    c = torch.tensor([ver], dtype=torch.float64, device=dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    return {"workload": f"configs[4]: Panda Quality, 256 restarts x {T} targets per GPU ({T * world} targets on {world} GPUs), "
                        "targets sharded, one NCCL all-gather of the per-target results",
            "value": float(c.item()) / (ms * 1e-3), "unit": "targets solved/s", "ms_per_call": ms,
            "success_rate": float(c.item()) / (T * world), "gathered_rows": int(q.shape[0]), "row_doubles": nq + 2}


# ----------------------------------------------------------------------------------------------- product arm
def run_product(args):
    import torch
    import torch.distributed as dist
    import optik_b200 as ob
    from optik_b200 import dist as obd

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = obd.bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned buffer exists
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)  # (NCCL's own output goes to stderr: main() re-routed fd 1)
    K, W, R = args.steps, max(args.warmup, 3), SEEDS_PER_GPU
    robot = ob.Robot.named(ROBOT)
    robot.set_device(local_rank)
    lb, ub = map(np.array, robot.joint_limits())
    n = robot.num_positions()
    tile = args.tile
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R, tol_f=TOL_F)
    x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev)
    lb_t, ub_t = torch.from_numpy(lb).to(dev), torch.from_numpy(ub).to(dev)
    rec = (torch.empty((R, n), dtype=torch.float64, device=dev), torch.empty((R,), dtype=torch.float64, device=dev),
           torch.empty((R,), dtype=torch.int32, device=dev), torch.empty((R,), dtype=torch.int32, device=dev))
    counters = torch.zeros(3, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    record = torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, device=dev)
    gathered = torch.empty((world, obd.RECORD_HEAD + n), dtype=torch.float64, device=dev)
    best = torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, device=dev)
    tally = torch.zeros(3, dtype=torch.int64, device=dev)  # verified, claimed, passes with a global best

    # cross-GPU best-pick: direct peer-to-peer stores over NVLink (two tiny kernels of ours); NCCL all-gather as fallback
    px = None
    if world > 1 and not args.nccl_exchange:
        px = obd.PeerExchange.create(robot, rank, world, device=dev)
        ok_all = torch.tensor([1.0 if px is not None else 0.0], device=dev)
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        if ok_all.item() < 1.0:
            px = None

    def one_pass(tgt):
        return obd.ik_restart_sharded(robot, cfg, tgt, x0, R, rank=rank, world=world, tile=tile, counters=counters, out=rec,
                                      record=record, gathered=gathered, best=best, exchange=px)

    def gate(tgt, best_rec):
        """success gate, OUTSIDE the timed region (between a pass's end event and the next start event), fully
        asynchronous: re-evaluate every record with the evaluator kernel and tally on the device"""
        q, f, st, ev = rec
        fe = robot.eval_batch(q, tgt, want=("f",))["f"]
        inside = ((q >= lb_t) & (q <= ub_t)).all(dim=1)
        claimed = st == 1
        tally[0] += (claimed & (fe < TOL_F) & inside).sum()
        tally[1] += claimed.sum()
        tally[2] += (best_rec[0] > 0).to(torch.int64)

    # ---- warm-up, and the size of a step: K steps must cover >= MIN_TIMED_SECONDS of device time
    wq = torch.from_numpy(seeded_q(lb, ub, W + 8, seed=41)).to(dev)
    wt = robot.eval_batch(wq, want=("ee",))["ee"].contiguous()
    for s in range(W):
        flush.fill_(s & 0xff)
        gate(wt[s], one_pass(wt[s])[0])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(8):
        one_pass(wt[W + s])
    b.record()
    torch.cuda.synchronize()
    est_ms = a.elapsed_time(b) / 8
    est = torch.tensor([est_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MIN)  # the same step size on every rank
    Pn = int(args.passes) if args.passes else max(1, int(np.ceil(MIN_TIMED_SECONDS * 1e3 / (K * float(est.item())))))
    NP = K * Pn
    qstar = seeded_q(lb, ub, NP)
    targets = robot.eval_batch(torch.from_numpy(qstar).to(dev), want=("ee",))["ee"].contiguous()  # resident in HBM
    counters.zero_()
    tally.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(NP)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(NP)]
    t_wall0 = time.perf_counter()
    for s in range(NP):
        flush.fill_(s & 0xff)
        ev0[s].record()
        best_rec, _ = one_pass(targets[s])
        ev1[s].record()
        gate(targets[s], best_rec)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    verified, claimed, found_passes = [int(x) for x in tally.cpu()]
    found_all = found_passes == NP
    cnt = counters.cpu().numpy().astype(np.int64)
    # oracle spot check of the last pass's records (golden-pinned evaluator, CPU)
    from oracle import oracle as O
    ch = O.Chain(robot.chain())
    q_h, st_h = rec[0].cpu().numpy(), rec[2].cpu().numpy()
    tgt_h = targets[NP - 1].cpu().numpy()
    idx = np.where(st_h == 1)[0][:: max(1, int((st_h == 1).sum()) // 512)][:512]
    oracle_ok = all(ch.objective(q_h[i], tgt_h) < TOL_F and np.all(q_h[i] >= lb) and np.all(q_h[i] <= ub) for i in idx)

    # ---- the same passes in throughput mode, still device-timed with device-resident inputs (single GPU only): DT_STREAMS
    # streams, each pass on SMs / E2E_SM_DIV blocks, so that several passes are co-resident.  This is the device-side
    # twin of the e2e leg below (it explains why e2e beats the isolated-pass `value`); converged attempts are the
    # kernel's own count (status == 1, which the gated loop above shows to equal the re-verified count).
    dev_tp = None
    if world == 1 and (tile or 1) == 1:
        DT_STREAMS, DT_PASSES = E2E_DEPTH_SINGLE, min(NP, 1024)
        dt_blocks = max(1, torch.cuda.get_device_properties(dev).multi_processor_count // E2E_SM_DIV)
        streams = [torch.cuda.Stream(device=dev) for _ in range(DT_STREAMS)]
        bufs = [((torch.empty((R, n), dtype=torch.float64, device=dev), torch.empty((R,), dtype=torch.float64, device=dev),
                  torch.empty((R,), dtype=torch.int32, device=dev), torch.empty((R,), dtype=torch.int32, device=dev)),
                 torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, device=dev)) for _ in range(DT_STREAMS)]
        cnt_tp = torch.zeros(3, dtype=torch.int64, device=dev)
        for k in range(DT_STREAMS):  # warm every stream
            with torch.cuda.stream(streams[k]):
                robot.ik_attempts(cfg, targets[k], x0, R, tile=tile, best=True, out=bufs[k][0], record=bufs[k][1], blocks=dt_blocks)
        torch.cuda.synchronize()
        flush.fill_(1)
        t_a = torch.cuda.Event(enable_timing=True)
        t_b = [torch.cuda.Event(enable_timing=True) for _ in range(DT_STREAMS)]
        torch.cuda.synchronize()
        t_a.record()
        for st_ in streams:
            st_.wait_event(t_a)
        for s in range(DT_PASSES):
            k = s % DT_STREAMS
            with torch.cuda.stream(streams[k]):
                robot.ik_attempts(cfg, targets[s], x0, R, tile=tile, best=True, counters=cnt_tp, out=bufs[k][0], record=bufs[k][1],
                                  blocks=dt_blocks)
        for k in range(DT_STREAMS):
            t_b[k].record(streams[k])
        torch.cuda.synchronize()
        tp_ms = max(t_a.elapsed_time(e) for e in t_b)
        conv_tp = int(cnt_tp[2])
        dev_tp = {"value": conv_tp / (tp_ms * 1e-3), "unit": UNIT, "ms_per_pass": tp_ms / DT_PASSES, "passes": DT_PASSES,
                  "streams": DT_STREAMS, "blocks_per_pass": dt_blocks, "evals_per_s": int(cnt_tp[1]) / (tp_ms * 1e-3),
                  "note": "device-resident inputs and outputs, CUDA events from the first launch to the last stream's end; "
                          "converged = status == 1 as counted by the kernel; L2 flushed once before the region"}
        del bufs, streams

    # ---- e2e: the same pass through the public host-buffer API (H2D of target/x0, D2H of every record, per pass).
    # Two legs: (a) one blocking call per pass; (b) the headline: calls enqueued on two library streams with two sets
    # of pinned buffers (OPTIK_BATCH_ASYNC), so one pass's D2H overlaps the next pass's kernels -- every pass still
    # copies its inputs from pinned host memory and its full records back, and the pass's result is read on the host.
    Ke = min(NP, int(os.environ.get("OPTIK_BENCH_E2E_PASSES", "2000")))
    tg_host = ob.pinned_empty((Ke, 8))
    tg_host[:] = targets[:Ke].cpu().numpy()
    x0_host = ob.pinned_empty(n)
    x0_host[:] = 0.5 * (lb + ub)
    sets = [((ob.pinned_empty((R, n)), ob.pinned_empty(R), ob.pinned_empty(R, np.int32), ob.pinned_empty(R, np.int32)),
             ob.pinned_empty(obd.RECORD_HEAD + n), ob.Stream(robot)) for _ in range(E2E_DEPTH_SINGLE)]

    def finish(k):
        """host side of a pass: wait for its stream, read the result (count converged records; cross-GPU best-pick)"""
        hrec, hrecord, stream = sets[k]
        stream.synchronize()
        if world > 1:
            robot.select_records(obd.all_gather_records(torch.from_numpy(hrecord).to(dev), out=gathered), out=best).cpu()
        return int((hrec[2] == 1).sum())

    # (a) blocking calls
    sync_conv, sync_times = 0, []
    Kb = min(Ke, 200)
    for s in range(3 + Kb):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        hrec, hrecord, stream = sets[0]
        robot.ik_attempts(cfg, tg_host[s % Ke], x0_host, R, restart_begin=rank * R, tile=tile, best=True, out=hrec,
                          record=hrecord, stream=stream, wait=False)
        c = finish(0)
        dt = time.perf_counter() - t0
        if s >= 3:
            sync_times.append(dt)
            sync_conv += c
    # (b) pipelined: the C ABI's asynchronous host-buffer call, E2E_DEPTH calls in flight.  Several GPUs: the same call
    # also stores its candidate record into every peer (fused push), and the best-pick select + the D2H of the global
    # best record are enqueued on the call's stream, so no rank blocks on the exchange between passes.
    if world > 1:
        dist.barrier()
    e2e_conv = 0
    host_submit_s = 0.0
    D = E2E_DEPTH_SINGLE if world == 1 else E2E_DEPTH_MULTI
    e2e_blocks = (max(1, torch.cuda.get_device_properties(dev).multi_processor_count // (E2E_SM_DIV if world == 1 else E2E_SM_DIV_MULTI))
                  if (tile or 1) == 1 else 0)
    if world == 1 or px is not None:
        while len(sets) < D:
            sets.append(((ob.pinned_empty((R, n)), ob.pinned_empty(R), ob.pinned_empty(R, np.int32), ob.pinned_empty(R, np.int32)),
                         ob.pinned_empty(obd.RECORD_HEAD + n), ob.Stream(robot)))
        ext = [torch.cuda.ExternalStream(s_[2].handle, device=dev) for s_ in sets] if world > 1 else None
        d_best = [torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, device=dev) for _ in sets] if world > 1 else None
        h_best = [torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, pin_memory=True) for _ in sets] if world > 1 else None

        def submit(s, k):
            hrec, hrecord, stream = sets[k]
            push = px.next_push() if world > 1 else None
            robot.ik_attempts(cfg, tg_host[s], x0_host, R, restart_begin=rank * R, tile=tile, best=True, out=hrec,
                              record=hrecord, stream=stream, wait=False, push=push, blocks=e2e_blocks)
            if world > 1:
                with torch.cuda.stream(ext[k]):
                    px.select(d_best[k], push[3])
                    h_best[k].copy_(d_best[k], non_blocking=True)

        def done(k):
            sets[k][2].synchronize()
            ok = 1 if world == 1 else int(h_best[k][0].item() >= 0)
            return int((sets[k][0][2] == 1).sum()) * ok

        for k in range(D):  # warm every slot
            submit(0, k)
        for k in range(D):
            done(k)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for s in range(Ke):
            k = s % D
            if s >= D:
                e2e_conv += done(k)
            ts = time.perf_counter()
            submit(s, k)
            host_submit_s += time.perf_counter() - ts
        for s in range(max(Ke - D, 0), Ke):
            e2e_conv += done(s % D)
    else:
        # NCCL fallback: optik_b200.dist.HostStepPipeline (torch copies + all-gather on a stream per slot), depth 4
        pipe = obd.HostStepPipeline(robot, cfg, R, rank=rank, world=world, tile=tile, depth=D, device=dev, exchange=px)
        for s in range(D):  # warm the slots (NCCL stream setup)
            pipe.submit(s, tg_host[0], x0_host)
        for s in range(D):
            pipe.result(s)
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(Ke):
            k = s % D
            if s >= D:
                r_ = pipe.result(k)
                e2e_conv += int((r_[2] == 1).sum()) * int(r_[4][0] >= 0)
            pipe.submit(k, tg_host[s], x0_host)
        for s in range(max(Ke - D, 0), Ke):
            r_ = pipe.result(s % D)
            e2e_conv += int((r_[2] == 1).sum()) * int(r_[4][0] >= 0)
    e2e_s = time.perf_counter() - t0
    h2d = 8 * 8 + n * 8
    d2h = R * (n * 8 + 8 + 4 + 4) + (obd.RECORD_HEAD + n) * 8

    # ---- roofline of the FK/Jacobian/error/gradient batch kernel (the path's HBM-bound kernel), measured live
    roof = None
    fp64_peak = None
    if rank == 0:
        B = 1 << 22
        rngq = torch.rand((B, n), dtype=torch.float64, device=dev) * (ub_t - lb_t) + lb_t
        tgB = robot.eval_batch(torch.rand((B, n), dtype=torch.float64, device=dev) * (ub_t - lb_t) + lb_t, want=("ee",))["ee"]
        times = []
        for i in range(2 + 5):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            outs = robot.eval_batch(rngq, tgB)
            b_.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(a.elapsed_time(b_))
            del outs
        ms = float(np.mean(times))
        bytes_per_eval = 8 * (8 * n + 17)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:  # DRAM bytes of this exact launch shape from the committed ncu --set full capture
            tf_name = "r02_eval_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_eval_traffic.json")) else "r01b_eval_traffic.json"
            tr = json.load(open(os.path.join(ROOT, "profiles", tf_name)))
            if tr["evals_per_launch"] == B:
                traffic = tr["traffic_bytes_per_launch"]
        except Exception:
            pass
        achieved = B * bytes_per_eval / (ms * 1e-3) / 1e9
        roof = {"kernel": "eval_kernel (batched FK + 6xn body Jacobian + se3-log error + gradient, fp64 I/O)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "traffic": traffic, "traffic_source": f"profiles/{tf_name} (ncu dram bytes read+write per launch)" if traffic else None,
                "algorithmic_bytes_per_launch": B * bytes_per_eval, "launch_ms": ms,
                "evals_per_launch": B, "inputs": "4 Mi configurations x (56 B q + 64 B target) in, 464 B out each: > L2"}
        # the reference's CPU path for the same work, timed beside it (oracle restatement of one objective callback with a
        # gradient request per configuration: FK once, Jacobian, Jlog6, log twice; all host cores and one core)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            chE = O.Chain(robot.chain())
            Bc = 1 << 20
            qh, th_ = rngq[:Bc].cpu().numpy(), tgB[:Bc].cpu().numpy()
            cores = os.cpu_count() or 1
            outc = O.eval_batch_threaded(chE, qh[:4096], th_[:4096], 1)
            gq = robot.eval_batch(rngq[:4096], tgB[:4096])
            ev_ok = bool(np.abs(gq["f"].cpu().numpy() - outc["f"]).max() <= 1e-12 and
                         np.abs(gq["jac"].cpu().numpy().reshape(4096, n, 6) - outc["jac"]).max() <= 1e-12)
            outc = O.eval_batch_threaded(chE, qh, th_, cores)  # first touch of the output pages
            t0 = time.perf_counter()
            O.eval_batch_threaded(chE, qh, th_, cores, out=outc)
            t_all = time.perf_counter() - t0
            t0 = time.perf_counter()
            O.eval_batch_threaded(chE, qh[:Bc // 8], th_[:Bc // 8], 1, out={k: v[:Bc // 8] for k, v in outc.items()})
            t_one = time.perf_counter() - t0
            roof["cpu_baseline"] = {"value": Bc / t_all, "unit": "configurations/s", "cores": cores, "kind": "port",
                                    "value_1_core": (Bc // 8) / t_one, "gpu_value": B / (ms * 1e-3),
                                    "gpu_matches_oracle_1e-12": ev_ok,
                                    "sample": f"{Bc} of the launch's Panda configurations, oracle/optik_oracle.c "
                                              "oracle_eval_batch_threaded (the work of one reference objective callback with "
                                              "gradient per configuration, lib.rs:305-337), fp64, gcc -O3; output pages touched first"}
            del outc
        del rngq, tgB
        # what HBM gives a kernel with this read/write mix and no arithmetic at all (120 B in, 464 B out per item -> 8 and
        # 29 sixteen-byte units, fully coalesced): the practical ceiling beside the copy figure
        mix = float(ob.load_library().optik_measure_hbm_mix(local_rank, B, 8, 29, 5))
        roof["mix_ceiling"] = {"achieved": mix, "unit": "GB/s", "frac_of_peak": mix / peak, "eval_frac_of_ceiling": achieved / mix if mix > 0 else None,
                               "what": "pure streaming kernel, 128 B read + 464 B written per item, same item count, coalesced, "
                                       "best of 5 launches (optik_measure_hbm_mix)"}
        fp64_peak = float(ob.load_library().optik_measure_fp64_peak(local_rank, 2.0))

    # ---- the other BASELINE configs and the per-target headline (every rank runs its own copy; rank 0 reports)
    with_cpu = world == 1 and rank == 0 and not args.no_cpu_baseline
    pt_cpu = cpu_per_target(ROBOT, 1 << 17, PER_TARGET_R, os.cpu_count() or 1) if with_cpu else None
    per_target = None
    configs = {}
    if not args.headline_only:
        per_target = per_target_block(ob, robot, dev, flush, rank, world, pt_cpu)
        if rank == 0:
            configs = other_configs(ob, robot, dev, flush, rank, world, with_cpu)
        if world > 1:
            dist.barrier()
            configs["config5"] = config5_block(ob, robot, dev, flush, rank, world)
            ptv = torch.tensor([per_target["value"], per_target["e2e"]["value"]], dtype=torch.float64, device=dev)
            dist.all_reduce(ptv, op=dist.ReduceOp.SUM)  # every rank solved its own 1 Mi targets concurrently-ish
            per_target["all_ranks_sum"] = {"value": float(ptv[0]), "e2e": float(ptv[1]),
                                           "note": "sum of the ranks' individually timed values (not barrier-aligned)"}

    # ---- reduce over ranks
    t = torch.tensor([dev_ms, e2e_s, t_wall, float(np.sum(sync_times))], dtype=torch.float64, device=dev)
    c = torch.tensor([verified, claimed, int(cnt[0]), int(cnt[1]), e2e_conv, sync_conv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s, t_wall, sync_s = [float(x) for x in t.cpu()]
    verified, claimed, attempts, evals, e2e_conv, sync_conv = [float(x) for x in c.cpu()]
    if rank == 0:
        cpu = None
        if with_cpu:
            threads = os.cpu_count() or 1
            cconv, csecs, catt = cpu_reference_passes(CPU_BASELINE_PASSES, 2, threads)
            cpu = {"value": cconv / csecs, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{CPU_BASELINE_PASSES} passes x {SEEDS_PER_GPU} seeds to one target each ({int(catt)} attempts, {csecs:.2f} s "
                             f"wall); oracle port of the rayon restart loop over the fp64 LM twin on all {threads} logical cores "
                             f"(the reference's default thread count, lib.rs:45); {CPU_FLAGS}; the Rust reference cannot be built here"}
        value = verified / (dev_ms * 1e-3)
        solve_bytes = 8 * (2 * n + 4)  # per attempt: q + f out, status/evals, target amortised
        # fp64 roofline of the solve kernel: executed fp64 instruction mix per objective evaluation from the committed ncu
        # capture (profiles/r02_solve_flops.json: DFMA, DMUL, DADD thread instructions / evaluations of that launch)
        flops = {"dfma": 899.0, "dmul": 453.0, "dadd": 166.0, "source": "builtin (r02b capture)"}
        try:
            flops = json.load(open(os.path.join(ROOT, "profiles", "r02_solve_flops.json")))
        except Exception:
            pass
        flop_per_eval = 2.0 * flops["dfma"] + flops["dmul"] + flops["dadd"]
        slots_per_eval = flops["dfma"] + flops["dmul"] + flops["dadd"]
        evals_per_s = evals / (dev_ms * 1e-3)
        ach_tf = evals_per_s * flop_per_eval / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world), "impl_detail": impl_detail(tile or 1, Pn),
            "ms_per_pass": dev_ms / NP, "passes": NP, "timed_device_seconds": dev_ms * 1e-3,
            "clocks": clocks,
            "e2e": {"value": e2e_conv / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d * Pn, "d2h_bytes_per_step": d2h * Pn,
                    "h2d_bytes_per_pass": h2d, "d2h_bytes_per_pass": d2h,
                    "passes": Ke, "ms_per_pass": e2e_s / Ke * 1e3, "pipeline_depth": E2E_DEPTH_SINGLE if world == 1 else E2E_DEPTH_MULTI,
                    "blocks_per_pass": e2e_blocks, "rank0_numa_node": numa_node,
                    "rank0_host_enqueue_ms_per_pass": host_submit_s / Ke * 1e3, "host_cores": os.cpu_count(),
                    "api": ("Robot.ik_attempts(pinned host buffers, stream=, wait=False) -> optik_gpu_ik_attempts with "
                            "OPTIK_BATCH_ASYNC (C ABI), one stream + buffer set per call in flight; host reads every pass's records"
                            if world == 1 else
                            "the same asynchronous host-buffer call with the candidate record stored into every peer by the solve "
                            "launch, then optik_gpu_exchange_select + D2H of the global best on the call's stream; %d passes in flight"
                            % E2E_DEPTH_MULTI if px is not None else
                            "optik_b200.dist.HostStepPipeline: per pass H2D from pinned buffers -> optik_gpu_ik_attempts "
                            "(device path) -> NCCL all-gather of the candidate record -> optik_gpu_select_records -> D2H of "
                            "the records and the global best, all on the pass's stream; %d passes in flight" % E2E_DEPTH_MULTI),
                    "blocking_call_value": sync_conv / sync_s, "blocking_call_ms_median": float(np.median(sync_times) * 1e3)},
            # per pass: solve_t1 (selection fused; the seed table is cached per robot) [+ exchange select | ncclAllGather + select_records]
            "gpu_launches": (1 if world == 1 else (2 if px is not None else 3)) * NP,
            "exchange": (None if world == 1 else "peer-to-peer stores over NVLink (optik_gpu_exchange_push/_select)" if px is not None
                         else "ncclAllGather + optik_gpu_select_records"),
            "device_throughput_mode": dev_tp,
            "roofline": roof,
            "roofline_solve": {"kernel": "solve_t1_kernel" if (tile or 1) == 1 else f"solve_kernel<{tile}>",
                               "bound": "fp64 issue / dependency latency (not HBM, not tensor)",
                               "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s (fp64)",
                               "frac": (ach_tf / fp64_peak) if fp64_peak and fp64_peak > 0 else None,
                               "peak_source": "measured in this run: optik_measure_fp64_peak (register-only DFMA kernel, 2 s)",
                               "flop_per_evaluation": flop_per_eval, "fp64_instructions_per_evaluation": slots_per_eval,
                               "fp64_issue_slot_frac": (evals_per_s * slots_per_eval * 2.0 / 1e12 / fp64_peak) if fp64_peak and fp64_peak > 0 else None,
                               "instruction_mix": flops,
                               "hbm_gbs": attempts * solve_bytes / (dev_ms * 1e-3) / 1e9,
                               "evals_per_s": evals_per_s, "attempts_per_s": attempts / (dev_ms * 1e-3)},
            "cpu_baseline": cpu,
            "per_target": per_target,
            "configs": configs,
            "success_rate_per_attempt": verified / max(attempts, 1.0),
            "verified_equals_claimed": verified == claimed, "oracle_spot_check_ok": bool(oracle_ok),
            "global_best_found_every_pass": bool(found_all),
            "wall_ms_per_pass_incl_flush_and_gate": t_wall / NP * 1e3,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL's version banner, library chatter)
    was re-routed to stderr by main() at the file-descriptor level."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--tile", type=int, default=0, help="lanes per restart seed: 8 (packed), 32 (one warp per seed); 0 = auto")
    ap.add_argument("--passes", type=int, default=0, help="passes per step; 0 = as many as make the timed region >= 1 s")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-exchange", action="store_true", help="use the NCCL all-gather for the cross-GPU best-pick")
    ap.add_argument("--headline-only", action="store_true", help="skip the per_target / configs blocks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
