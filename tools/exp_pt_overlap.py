"""Per-target batches (1 Mi Panda targets, Speed) in throughput mode: calls on several streams, each on a fraction of
the machine, device-resident inputs: does overlapping the batches' tails pay as it does for the 65 536-seed passes?"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import optik_b200 as ob
r = ob.Robot.named("panda")
n = 7
T = 1 << 20
lb, ub = [torch.tensor(x, dtype=torch.float64, device="cuda") for x in r.joint_limits()]
g = torch.Generator(device="cuda").manual_seed(42)
qs = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb
x0 = (torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * (ub - lb) + lb).contiguous()
tg = r.eval_batch(qs, want=("ee",))["ee"].contiguous()
cfg = ob.SolverConfig(solution_mode="speed", max_time=0.0, max_restarts=32)
for blocks, S in ((0, 1), (0, 2), (0, 3), (222, 2), (148, 2), (148, 3), (148, 4), (99, 4), (74, 4), (74, 6)):
    streams = [torch.cuda.Stream() for _ in range(S)]
    outs = [None] * S
    for k in range(S):
        with torch.cuda.stream(streams[k]):
            outs[k] = r.ik_batch(cfg, tg, x0, restarts=32, blocks=blocks)
    torch.cuda.synchronize()
    N = 12
    a = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    a.record()
    for st in streams: st.wait_event(a)
    for i in range(N):
        with torch.cuda.stream(streams[i % S]):
            outs[i % S] = r.ik_batch(cfg, tg, x0, restarts=32, blocks=blocks)
    for k in range(S): ends[k].record(streams[k])
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(e) for e in ends) / N
    ok = float(cfg.is_success(outs[0][2].cpu().numpy()).mean())
    print(f"blocks {blocks:3d} streams {S}: {ms:.3f} ms/call  {T * ok / ms * 1e3:.3e} targets/s", flush=True)
