#!/usr/bin/env python
"""bench.py -- headline benchmark of the IK hot path (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, all host threads)

Workload ("step"): Panda 7-DOF, 65 536 random-restart seeds to ONE reachable target per GPU, SolutionMode::Quality
(every restart runs, lib.rs:398-407), tol_f = 1e-6, max_time = 0.  Rank g runs the restart range
[g*65536, (g+1)*65536) (weak scaling); each rank selects its best candidate on the device and ONE NCCL all-gather of an
11-double record per rank picks the global best.  A fresh target is used every step.

Metric: IK solves/s, success-gated = restart attempts whose solution, RE-EVALUATED by the evaluator kernel outside the
timed region, satisfies the reference's success predicate f(q) < tol_f with lb <= q <= ub (and a sample of which is
re-checked by the golden-pinned CPU oracle), divided by the device time of the steps (CUDA events, max over ranks).
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
CPU_BASELINE_STEPS = 32  # cpu_baseline sample: 32 full steps (2 Mi attempts, ~20 CPU-seconds on 16 cores)
E2E_DEPTH_SINGLE = int(os.environ.get("OPTIK_BENCH_E2E_DEPTH", "2"))  # host-buffer calls in flight on one GPU
E2E_DEPTH_MULTI = 4  # host-buffer steps in flight per rank when a collective sits inside the step (N > 1)
ROBOT = "panda"
TOL_F = 1e-6
METRIC = "IK solves/sec (success-gated, Panda 7-DOF)"
UNIT = "solves/s"


def workload_config(n_gpus, tile):
    return {
        "workload": "configs[1]: Panda 7-DOF, 65536 random-restart seeds to one target per GPU per step, "
                    "SolutionMode::Quality, tol_f=1e-6, max_time=0, fresh reachable target every step",
        "robot": ROBOT, "dof": 7, "seeds_per_gpu_per_step": SEEDS_PER_GPU, "targets_per_step": 1,
        "lanes_per_seed": tile, "layout": ("thread-per-seed kernel (solve_t1_kernel)" if tile == 1 else
                                           f"tile kernel, {tile} lanes per seed (solve_kernel<{tile}>)" if tile else "cpu"),
        "solve_definition": "one restart attempt that converged (f<tol_f inside the joint limits, re-verified)",
        "l2": "flushed (256 MiB write) before every timed step; steps timed individually with CUDA events",
        "parallelism": f"restart-range sharding over {n_gpus} GPU(s); one NCCL all-gather of 11 doubles/rank/step "
                       "for the Quality best-pick" if n_gpus > 1 else "single GPU",
    }


def make_targets(count, seed=42):
    """Seeded joint vectors uniform in the limits -> reachable targets (examples/example.rs:24-26 protocol)."""
    import optik_b200 as ob
    r = ob.Robot.named(ROBOT)
    lb, ub = map(np.array, r.joint_limits())
    rng = np.random.default_rng(seed)
    return r, lb, ub, rng.uniform(lb, ub, size=(count, len(lb)))


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max, self.error = [], set(), None, None
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
             "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.error:
            d["error"] = self.error
        return d


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_steps(steps, warmup, threads, seeds=SEEDS_PER_GPU):
    """The reference's CPU path for this workload: the oracle port of the rayon restart loop (oracle/ref_loop.c over the
    fp64 solver twin) with `threads` host threads.  Returns (converged, seconds, attempts)."""
    from oracle import oracle as O
    import optik_b200 as ob
    r, lb, ub, qstar = make_targets(steps + warmup)
    ch = O.Chain(r.chain())
    x0 = 0.5 * (lb + ub)
    conv = att = 0
    total = 0.0
    for s in range(steps + warmup):
        tgt = ch.fk(qstar[s])[1]
        t0 = time.perf_counter()
        res = O.ref_ik_threaded(ch, tgt, x0, 0, seeds, "quality", threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            conv += res["converged"]
            att += res["attempts"]
            total += dt
    return conv, total, att


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = args.steps
    conv, secs, att = cpu_reference_steps(steps, args.warmup, threads)
    value = conv / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, 0),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x {SEEDS_PER_GPU} seeds to one target each (the full per-GPU step), "
                                   "oracle port of the rayon restart loop over the fp64 LM twin; the Rust reference "
                                   "(NLopt SLSQP) cannot be built in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "attempts_per_s": att / secs, "gpu_launches": 0,
    }
    emit(line)


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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    K, W, R = args.steps, max(args.warmup, 3), SEEDS_PER_GPU
    robot, lb, ub, qstar = make_targets(K + W)
    robot.set_device(local_rank)
    n = robot.num_positions()
    tile = args.tile
    cfg = ob.SolverConfig(solution_mode="quality", max_time=0.0, max_restarts=R, tol_f=TOL_F)
    # inputs resident in HBM before the timed region
    targets = robot.eval_batch(torch.from_numpy(qstar).to(dev), want=("ee",))["ee"].contiguous()
    x0 = torch.from_numpy(0.5 * (lb + ub)).to(dev)
    lb_t, ub_t = torch.from_numpy(lb).to(dev), torch.from_numpy(ub).to(dev)
    rec = (torch.empty((R, n), dtype=torch.float64, device=dev), torch.empty((R,), dtype=torch.float64, device=dev),
           torch.empty((R,), dtype=torch.int32, device=dev), torch.empty((R,), dtype=torch.int32, device=dev))
    counters = torch.zeros(3, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    record = torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, device=dev)
    gathered = torch.empty((world, obd.RECORD_HEAD + n), dtype=torch.float64, device=dev)
    best = torch.empty((obd.RECORD_HEAD + n,), dtype=torch.float64, device=dev)
    tally = torch.zeros(3, dtype=torch.int64, device=dev)  # verified, claimed, steps with a global best

    def step(s):
        return obd.ik_restart_sharded(robot, cfg, targets[s], x0, R, rank=rank, world=world, tile=tile,
                                      counters=counters, out=rec, record=record, gathered=gathered, best=best)

    def gate(s, best_rec):
        """success gate, OUTSIDE the timed region (between the step's end event and the next start event), fully
        asynchronous: re-evaluate every record with the evaluator kernel and tally on the device"""
        q, f, st, ev = rec
        fe = robot.eval_batch(q, targets[s], want=("f",))["f"]
        inside = ((q >= lb_t) & (q <= ub_t)).all(dim=1)
        claimed = st == 1
        tally[0] += (claimed & (fe < TOL_F) & inside).sum()
        tally[1] += claimed.sum()
        tally[2] += (best_rec[0] > 0).to(torch.int64)

    for s in range(W):
        flush.fill_(s & 0xff)
        gate(s, step(s)[0])
    torch.cuda.synchronize()
    counters.zero_()
    tally.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for s in range(K):
        flush.fill_(s & 0xff)
        ev0[s].record()
        best_rec, _ = step(W + s)
        ev1[s].record()
        gate(W + s, best_rec)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    verified, claimed, found_steps = [int(x) for x in tally.cpu()]
    found_all = found_steps == K
    cnt = counters.cpu().numpy().astype(np.int64)
    # oracle spot check of the last step's records (golden-pinned evaluator, CPU)
    from oracle import oracle as O
    ch = O.Chain(robot.chain())
    q_h, st_h = rec[0].cpu().numpy(), rec[2].cpu().numpy()
    tgt_h = targets[W + K - 1].cpu().numpy()
    idx = np.where(st_h == 1)[0][:: max(1, int((st_h == 1).sum()) // 512)][:512]
    oracle_ok = all(ch.objective(q_h[i], tgt_h) < TOL_F and np.all(q_h[i] >= lb) and np.all(q_h[i] <= ub) for i in idx)

    # ---- e2e: the same step through the public host-buffer API (H2D of target/x0, D2H of every record, per step).
    # Two legs: (a) one blocking call per step; (b) the headline: calls enqueued on two library streams with two sets
    # of pinned buffers (OPTIK_BATCH_ASYNC), so one step's D2H overlaps the next step's kernels -- every step still
    # copies its inputs from pinned host memory and its full records back, and the step's result is read on the host.
    tg_host = ob.pinned_empty((K + W, 8))
    tg_host[:] = targets.cpu().numpy()
    x0_host = ob.pinned_empty(n)
    x0_host[:] = 0.5 * (lb + ub)
    Ke = min(K, int(os.environ.get("OPTIK_BENCH_E2E_STEPS", "200")))
    sets = [((ob.pinned_empty((R, n)), ob.pinned_empty(R), ob.pinned_empty(R, np.int32), ob.pinned_empty(R, np.int32)),
             ob.pinned_empty(obd.RECORD_HEAD + n), ob.Stream(robot)) for _ in range(E2E_DEPTH_SINGLE)]

    def finish(k):
        """host side of a step: wait for its stream, read the result (count converged records; cross-GPU best-pick)"""
        hrec, hrecord, stream = sets[k]
        stream.synchronize()
        if world > 1:
            robot.select_records(obd.all_gather_records(torch.from_numpy(hrecord).to(dev), out=gathered), out=best).cpu()
        return int((hrec[2] == 1).sum())

    # (a) blocking calls
    sync_conv, sync_times = 0, []
    for s in range(3 + Ke):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        hrec, hrecord, stream = sets[0]
        robot.ik_attempts(cfg, tg_host[W + (s % K)], x0_host, R, restart_begin=rank * R, tile=tile, best=True, out=hrec,
                          record=hrecord, stream=stream, wait=False)
        c = finish(0)
        dt = time.perf_counter() - t0
        if s >= 3:
            sync_times.append(dt)
            sync_conv += c
    # (b) pipelined, depth 2.  One GPU: the C ABI's asynchronous host-buffer call.  Several GPUs: the same step with the
    # all-gather + global best-pick enqueued on the step's stream (optik_b200.dist.HostStepPipeline), so no rank blocks
    # on a collective between steps.
    if world > 1:
        dist.barrier()
    e2e_conv = 0
    if world == 1:
        t0 = time.perf_counter()
        for s in range(Ke):
            k = s % E2E_DEPTH_SINGLE
            if s >= E2E_DEPTH_SINGLE:
                e2e_conv += finish(k)
            hrec, hrecord, stream = sets[k]
            robot.ik_attempts(cfg, tg_host[W + (s % K)], x0_host, R, restart_begin=rank * R, tile=tile, best=True, out=hrec,
                              record=hrecord, stream=stream, wait=False)
        for s in range(max(Ke - E2E_DEPTH_SINGLE, 0), Ke):
            e2e_conv += finish(s % E2E_DEPTH_SINGLE)
    else:
        # depth 4: the NCCL kernel of step s only gets SM resources once step s+1's solve kernel drains, so a step
        # completes about one step late; four slots keep two solve kernels in flight regardless
        D = E2E_DEPTH_MULTI
        pipe = obd.HostStepPipeline(robot, cfg, R, rank=rank, world=world, tile=tile, depth=D, device=dev)
        for s in range(D):  # warm the slots (NCCL stream setup)
            pipe.submit(s, tg_host[W], x0_host)
        for s in range(D):
            pipe.result(s)
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(Ke):
            k = s % D
            if s >= D:
                r_ = pipe.result(k)
                e2e_conv += int((r_[2] == 1).sum()) * int(r_[4][0] >= 0)
            pipe.submit(k, tg_host[W + (s % K)], x0_host)
        for s in range(max(Ke - D, 0), Ke):
            r_ = pipe.result(s % D)
            e2e_conv += int((r_[2] == 1).sum()) * int(r_[4][0] >= 0)
    e2e_s = time.perf_counter() - t0
    h2d = 8 * 8 + n * 8
    d2h = R * (n * 8 + 8 + 4 + 4) + (obd.RECORD_HEAD + n) * 8

    # ---- roofline of the FK/Jacobian/error/gradient batch kernel (the path's HBM-bound kernel), measured live
    roof = None
    if rank == 0:
        B = 1 << 22
        rngq = torch.rand((B, n), dtype=torch.float64, device=dev) * (ub_t - lb_t) + lb_t
        tgB = robot.eval_batch(torch.rand((B, n), dtype=torch.float64, device=dev) * (ub_t - lb_t) + lb_t, want=("ee",))["ee"]
        outs = None
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
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01b_eval_traffic.json")))
            if tr["evals_per_launch"] == B:
                traffic = tr["traffic_bytes_per_launch"]
        except Exception:
            pass
        achieved = B * bytes_per_eval / (ms * 1e-3) / 1e9
        roof = {"kernel": "eval_kernel (batched FK + 6xn body Jacobian + se3-log error + gradient, fp64 I/O)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "traffic": traffic, "traffic_source": "profiles/r01b_eval_traffic.json (ncu dram bytes read+write per launch)",
                "algorithmic_bytes_per_launch": B * bytes_per_eval, "launch_ms": ms,
                "evals_per_launch": B, "inputs": "4 Mi configurations x (56 B q + 64 B target) in, 464 B out each: > L2"}
        del rngq, tgB

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
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cconv, csecs, catt = cpu_reference_steps(CPU_BASELINE_STEPS, 2, threads)
            cpu = {"value": cconv / csecs, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{CPU_BASELINE_STEPS} steps x {SEEDS_PER_GPU} seeds to one target each ({int(catt)} attempts, {csecs:.2f} s wall); "
                             "oracle port of the rayon restart loop over the fp64 LM twin (the Rust reference cannot be built here)"}
        value = verified / (dev_ms * 1e-3)
        solve_bytes = 8 * (2 * n + 4)  # per attempt: q + f out, status/evals, target amortised
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, tile or 1),
            "clocks": clocks,
            "e2e": {"value": e2e_conv / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "ms_per_step": e2e_s / Ke * 1e3, "pipeline_depth": E2E_DEPTH_SINGLE if world == 1 else E2E_DEPTH_MULTI,
                    "api": ("Robot.ik_attempts(pinned host buffers, stream=, wait=False) -> optik_gpu_ik_attempts with "
                            "OPTIK_BATCH_ASYNC (C ABI), one stream + buffer set per call in flight; host reads every step's records"
                            if world == 1 else
                            "optik_b200.dist.HostStepPipeline: per step H2D from pinned buffers -> optik_gpu_ik_attempts "
                            "(device path) -> NCCL all-gather of the candidate record -> optik_gpu_select_records -> D2H of "
                            "the records and the global best, all on the step's stream; %d steps in flight" % E2E_DEPTH_MULTI),
                    "blocking_call_value": sync_conv / sync_s, "blocking_call_ms_median": float(np.median(sync_times) * 1e3)},
            "gpu_launches": (3 if world == 1 else 4) * K,  # per step: solve_t1 + select (slice pass + final pass) [+ select_records]
            "roofline": roof,
            "roofline_solve": {"kernel": "solve_t1_kernel" if (tile or 1) == 1 else f"solve_kernel<{tile}>", "bound": "fp64 issue / latency (not HBM)",
                               "hbm_gbs": attempts * solve_bytes / (dev_ms * 1e-3) / 1e9,
                               "evals_per_s": evals / (dev_ms * 1e-3), "attempts_per_s": attempts / (dev_ms * 1e-3)},
            "cpu_baseline": cpu,
            "success_rate_per_attempt": verified / max(attempts, 1.0),
            "verified_equals_claimed": verified == claimed, "oracle_spot_check_ok": bool(oracle_ok),
            "global_best_found_every_step": bool(found_all),
            "wall_ms_per_step_incl_flush_and_gate": t_wall / K * 1e3,
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--tile", type=int, default=0, help="lanes per restart seed: 8 (packed), 32 (one warp per seed); 0 = auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
