"""Multi-GPU layer: one process per GPU, sharding with no data-path collective except the best-pick exchange.

The reference has no distributed backend (single-process rayon); what shards here is its restart
index space (crates/optik/src/lib.rs:298-301) and, for batches, the independent targets.

  * restart sharding (BASELINE config 2, single-target calls): rank g runs restarts [g*R, (g+1)*R); every rank
    keeps its best candidate; ONE all-gather of a (4+n)-double record per rank; every rank then applies the
    reference's selection rule (lib.rs:397-413) to the gathered records -- Quality: arg-min ||q-x0||^2,
    Speed: lowest converged restart index.
  * target sharding (configs 3, 5): contiguous target ranges; selection is rank-local; one all-gather assembles
    the result (optional).

torch.distributed is plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous [begin, end) of `total` items for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


RECORD_HEAD = 8  # [found, score, restart, cost, status, 0, 0, 0, q...]  (include/optik_b200.h)


def pack_candidate(found, score, restart, cost, q, status=1.0):
    """-> float64 candidate record (torch tensor on q's device) in the layout the select kernels use."""
    import torch
    f64 = lambda v: torch.as_tensor(v, dtype=torch.float64, device=q.device).reshape(())
    head = torch.stack([f64(found), f64(score), f64(restart), f64(cost), f64(status), f64(0.0), f64(0.0), f64(0.0)])
    return torch.cat([head, q.to(torch.float64).reshape(-1)])


def select_candidates(records, as_tensor=False):
    """Reference selection (lib.rs:397-413) over gathered records (W, RECORD_HEAD+n) with torch ops (CPU tests, and
    the specification of optik_gpu_select_records): converged first, then lowest score, then lowest restart index.
    Returns (index, record); with as_tensor=True only the record, without a host sync."""
    import torch
    found, score, restart = records[:, 0], records[:, 1], records[:, 2]
    big = torch.finfo(torch.float64).max
    key_score = torch.where(found > 0, score, torch.full_like(score, big))
    best = key_score.min()
    tie = (key_score == best) & ((found > 0) | (found.max() <= 0))
    r = torch.where(tie, restart, torch.full_like(restart, big))
    if as_tensor:
        return records.index_select(0, torch.argmin(r).reshape(1))[0]
    idx = int(torch.argmin(r))
    return idx, records[idx]


def all_gather_records(record, group=None, out=None):
    """One all-gather of a small fixed-size record per rank -> (world, len) tensor.  NCCL for CUDA tensors."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return record.unsqueeze(0)
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world,) + tuple(record.shape), dtype=record.dtype, device=record.device)
    if record.is_cuda:
        dist.all_gather_into_tensor(out, record.contiguous(), group=group)  # NCCL over NVLink / NVSwitch
    else:
        dist.all_gather(list(out.unbind(0)), record.contiguous(), group=group)  # gloo (CPU tests)
    return out


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and the node's
    cpulist), before pinned host buffers are allocated: they then live in the memory next to the GPU's PCIe root, and
    eight ranks' D2H streams do not all cross to one socket.  Returns the node, or None when nothing was changed
    (single node, unknown topology, cpuset without those CPUs)."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus, dom, dv = pr.pci_bus_id, pr.pci_domain_id, pr.pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dv:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0 or not os.path.isdir(f"/sys/devices/system/node/node{node}"):
            return None
        if len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]) < 2:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class PeerExchange:
    """The best-pick exchange as direct peer-to-peer stores over NVLink (optik_gpu_exchange_push / _select,
    csrc/exchange_kernel.cu) instead of an NCCL all-gather: every rank owns one symmetric buffer that torch's symmetric
    memory maps into every peer (plumbing: allocation + handle exchange only; the data path is our two kernels).
    `PeerExchange.create` returns None where peer mapping is unavailable (callers then use the NCCL all-gather)."""

    def __init__(self, robot, rank, world, group, device):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import load_library, _check
        self._lib, self._check, self.robot, self.rank, self.world = load_library(), _check, robot, rank, world
        nbytes = int(self._lib.optik_gpu_exchange_bytes(robot._h, world))
        self.buf = symm.empty(nbytes // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        grp = group if group is not None else dist.group.WORLD
        self.handle = symm.rendezvous(self.buf, grp.group_name if hasattr(grp, "group_name") else grp)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == world and ptrs[rank] == self.buf.data_ptr()
        self.peers = torch.tensor(ptrs, dtype=torch.int64, device=device)  # device array of the peers' base addresses
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # every buffer is zeroed before anybody pushes

    @staticmethod
    def create(robot, rank, world, group=None, device=None):
        import torch
        try:
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            return PeerExchange(robot, rank, world, group, dev)
        except Exception as e:  # no peer access / symmetric memory on this system
            import sys
            print(f"optik_b200.dist: peer-to-peer exchange unavailable ({e!r}); using the NCCL all-gather", file=sys.stderr)
            return None

    def exchange(self, record, out):
        """Enqueue push + select on the current stream: `record` (this rank's candidate) -> `out` (the global best)."""
        import torch
        self.seq += 1
        stream = torch.cuda.current_stream(record.device).cuda_stream
        self._check(self._lib.optik_gpu_exchange_push(self.robot._h, record.data_ptr(), self.peers.data_ptr(), self.rank,
                                                      self.world, self.seq, stream))
        return self.select(out, self.seq)

    def next_push(self):
        """Arguments for a solve launch that pushes its candidate record itself (Robot.ik_attempts(push=...))."""
        self.seq += 1
        return (self.peers.data_ptr(), self.rank, self.world, self.seq)

    def select(self, out, seq):
        import torch
        stream = torch.cuda.current_stream(out.device).cuda_stream
        self._check(self._lib.optik_gpu_exchange_select(self.robot._h, self.buf.data_ptr(), self.world, int(seq),
                                                        out.data_ptr(), stream))
        return out


def ik_restart_sharded(robot, config, target, x0, restarts_per_rank, rank=0, world=1, group=None, tile=0,
                       counters=None, out=None, record=None, gathered=None, best=None, exchange=None):
    """One target, restarts sharded over ranks: rank g runs [g*R, (g+1)*R), selects its best candidate on the
    device (select_kernel), ONE all-gather of the candidate record, then the same selection rule over the gathered
    records (optik_gpu_select_records) gives every rank the GLOBAL best under config.solution_mode.  With
    `exchange` (a PeerExchange) the all-gather is replaced by direct peer-to-peer stores + a flag wait.
    Returns (best_record, local_records); everything stays on the device, no host sync.
    target (8,), x0 (n,) CUDA float64 tensors; record/gathered/best are optional preallocated buffers."""
    R = int(restarts_per_rank)
    fused_push = exchange is not None and world > 1 and tile in (0, 1) and robot.num_positions() <= 8
    push = exchange.next_push() if fused_push else None
    q, f, st, ev, rec = robot.ik_attempts(config, target, x0, R, restart_begin=rank * R, tile=tile, best=True,
                                          counters=counters, out=out, record=record, push=push)
    if world == 1:
        return rec, (q, f, st, ev)
    if exchange is not None:  # direct peer stores over NVLink (by the solve launch itself when it can) + a flag wait
        import torch
        if best is None:
            best = torch.empty_like(rec)
        if fused_push:
            return exchange.select(best, push[3]), (q, f, st, ev)
        return exchange.exchange(rec, best), (q, f, st, ev)
    allrec = all_gather_records(rec, group, out=gathered)
    return robot.select_records(allrec, out=best), (q, f, st, ev)


def ik_batch_target_sharded(robot, config, targets, x0, restarts, rank=0, world=1, group=None, gather=True, **kw):
    """T targets sharded contiguously over ranks (selection is local).  targets/x0 are this rank's CUDA tensors for
    its shard; with gather=True the per-target results are assembled on every rank with one all-gather."""
    import torch
    import torch.distributed as dist
    q, f, st = robot.ik_batch(config, targets, x0, restarts=restarts, **kw)
    if not gather or world == 1 or not (dist.is_available() and dist.is_initialized()):
        return q, f, st
    n = q.shape[1]
    rec = torch.cat([q, f[:, None], st.to(torch.float64)[:, None]], dim=1).contiguous()
    out = torch.empty((world * rec.shape[0], n + 2), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)  # equal shard sizes required
    return out[:, :n], out[:, n], out[:, n + 1].to(torch.int32)


class HostStepPipeline:
    """Host-buffer steps of ik_restart_sharded with the cross-GPU best-pick INSIDE the step, `depth` steps in flight.

    Each slot owns a CUDA stream, device staging and pinned host buffers; submit() enqueues, on the slot's stream:
    H2D of (target, x0) from pinned memory -> solve + select (C ABI, device path) -> one NCCL all-gather of the
    candidate record -> select_records -> D2H of this rank's per-restart records and of the global best record.
    Nothing blocks the host until result(slot).  With world == 1 the all-gather is skipped."""

    def __init__(self, robot, config, restarts_per_rank, rank=0, world=1, group=None, tile=0, depth=2, device=None,
                 exchange=None):
        import torch
        assert exchange is None or depth <= 16, "the peer exchange keeps at most 16 calls in flight"
        self.exchange = exchange
        self.robot, self.config, self.R, self.rank, self.world, self.group, self.tile = robot, config, int(restarts_per_rank), rank, world, group, tile
        n = robot.num_positions()
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.slots = []
        for _ in range(depth):
            pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
            d = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
            self.slots.append({
                "stream": torch.cuda.Stream(device=dev),
                "h_in": (pin((8,), torch.float64), pin((n,), torch.float64)),
                "d_in": (d((8,), torch.float64), d((n,), torch.float64)),
                "d_out": (d((self.R, n), torch.float64), d((self.R,), torch.float64), d((self.R,), torch.int32), d((self.R,), torch.int32)),
                "h_out": (pin((self.R, n), torch.float64), pin((self.R,), torch.float64), pin((self.R,), torch.int32), pin((self.R,), torch.int32)),
                "record": d((RECORD_HEAD + n,), torch.float64), "gathered": d((world, RECORD_HEAD + n), torch.float64),
                "best": d((RECORD_HEAD + n,), torch.float64), "h_best": pin((RECORD_HEAD + n,), torch.float64),
                "busy": False,
            })

    def submit(self, slot, target, x0):
        """target (8,), x0 (n,): host arrays; copied into the slot's pinned buffers, then everything is enqueued."""
        import torch
        s = self.slots[slot]
        if s["busy"]:  # the slot's pinned buffers are still owned by its previous step
            s["stream"].synchronize()
        s["busy"] = True
        s["h_in"][0].numpy()[:] = target
        s["h_in"][1].numpy()[:] = x0
        with torch.cuda.stream(s["stream"]):
            s["d_in"][0].copy_(s["h_in"][0], non_blocking=True)
            s["d_in"][1].copy_(s["h_in"][1], non_blocking=True)
            best, _ = ik_restart_sharded(self.robot, self.config, s["d_in"][0], s["d_in"][1], self.R, rank=self.rank,
                                         world=self.world, group=self.group, tile=self.tile, out=s["d_out"],
                                         record=s["record"], gathered=s["gathered"], best=s["best"], exchange=self.exchange)
            for h, dv in zip(s["h_out"], s["d_out"]):
                h.copy_(dv, non_blocking=True)
            s["h_best"].copy_(best, non_blocking=True)

    def result(self, slot):
        """Waits for the slot's step; returns (q, f, status, evals, global_best_record) as numpy views of pinned memory."""
        s = self.slots[slot]
        s["stream"].synchronize()
        s["busy"] = False
        return tuple(h.numpy() for h in s["h_out"]) + (s["h_best"].numpy(),)
