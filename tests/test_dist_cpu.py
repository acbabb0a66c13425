"""CPU: the N>1 host logic (sharding + best-pick exchange) with world_size-2 gloo processes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optik_b200 import dist as obd


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [obd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_select_candidates_rules():  # lib.rs:397-413
    rec = torch.tensor([[0.0, 0.1, 3, 9.0, 5, 0, 0, 0, 1, 1],      # not converged: never wins while a converged one exists
                        [1.0, 2.5, 7, 1e-7, 1, 0, 0, 0, 2, 2],
                        [1.0, 0.5, 9, 2e-7, 1, 0, 0, 0, 3, 3],      # lowest score
                        [1.0, 0.5, 4, 3e-7, 1, 0, 0, 0, 4, 4]],     # same score, lower restart index -> wins
                       dtype=torch.float64)
    idx, best = obd.select_candidates(rec)
    assert idx == 3 and best[2] == 4
    assert torch.equal(obd.select_candidates(rec, as_tensor=True), rec[3])
    none = rec.clone()
    none[:, 0] = 0
    idx, best = obd.select_candidates(none)
    assert best[0] == 0


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank holds the best candidate of its restart range; the exchange must give every rank the global best
        rng = np.random.default_rng(100 + rank)
        n = 7
        local = []
        for trial in range(20):
            found = torch.tensor(float(rng.random() < 0.8), dtype=torch.float64)
            score = torch.tensor(rng.random(), dtype=torch.float64)
            restart = torch.tensor(float(rank * 1000 + trial), dtype=torch.float64)
            rec = obd.pack_candidate(found, score, restart, torch.tensor(1e-7, dtype=torch.float64),
                                     torch.from_numpy(rng.random(n)))
            allrec = obd.all_gather_records(rec)
            assert allrec.shape == (world, obd.RECORD_HEAD + n)
            assert torch.equal(allrec[rank], rec)
            idx, best = obd.select_candidates(allrec)
            local.append(best.numpy().copy())
        np.save(os.path.join(tmp, f"best_{rank}.npy"), np.array(local))
    finally:
        dist.destroy_process_group()


def test_best_pick_exchange_world2_gloo(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "best_0.npy"), np.load(tmp_path / "best_1.npy")
    assert np.array_equal(a, b)  # every rank arrives at the same global best
    # and it is the arg-min over both ranks' candidates
    for trial in range(20):
        cands = []
        for rank in range(world):
            rng = np.random.default_rng(100 + rank)
            for t in range(trial + 1):
                found, score = float(rng.random() < 0.8), rng.random()
                q = rng.random(7)
            cands.append((found, score, rank * 1000 + trial))
        conv = [c for c in cands if c[0] > 0]
        if conv:
            want = min(conv, key=lambda c: (c[1], c[2]))
            assert a[trial][0] == 1.0 and a[trial][2] == want[2]
        else:
            assert a[trial][0] == 0.0


def test_numa_binding_is_a_no_op_without_topology():
    """bind_to_gpu_numa_node never raises and leaves the affinity alone when it cannot resolve the GPU's node (no CUDA
    device here; single-node hosts)."""
    import os
    from optik_b200 import dist as obd
    before = os.sched_getaffinity(0)
    assert obd.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
