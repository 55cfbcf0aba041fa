"""The N>1 host logic on CPU: world_size-2 gloo processes shard segments by length-balanced bucketing, each computes its
rows, rank 0 gathers them back into global order (same code path bench.py uses with NCCL)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_partition_is_balanced_and_complete():
    sys.path.insert(0, str(ROOT))
    from prosody_b200 import shard
    rng = np.random.default_rng(0)
    costs = np.concatenate([rng.uniform(3, 12, 600), rng.uniform(10, 30, 300), rng.uniform(20, 40, 100)])   # C5-like length mix
    for world in (2, 4, 8):
        parts = shard.partition_by_cost(costs, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(costs)))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.01
        assert all(p == sorted(p) for p in parts)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from prosody_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        n_seg = 37
        rows_per_seg = rng.integers(1, 9, n_seg)
        costs = rng.uniform(1, 10, n_seg)
        first = np.concatenate([[0], np.cumsum(rows_per_seg)])
        parts = shard.partition_by_cost(costs, world)
        mine = parts[rank]
        ids = np.concatenate([np.arange(first[s], first[s + 1]) for s in mine]) if mine else np.zeros(0, np.int64)
        rows = np.stack([ids * 1.5, np.sin(ids), ids % 7, -ids, ids * ids], 1).astype(np.float64) if len(ids) else np.zeros((0, 5))
        out = shard.gather_rows(torch.from_numpy(rows), torch.from_numpy(ids), dst=0)
        sizes = shard.row_counts(len(ids), torch.device("cpu"))              # exchanged once, reused by later gathers
        assert sum(sizes) == int(first[-1]) and sizes[rank] == len(ids)
        again = shard.gather_rows(torch.from_numpy(rows), torch.from_numpy(ids), dst=0, sizes=sizes)
        assert (again is None) == (rank != 0) and (rank != 0 or np.array_equal(again.numpy(), out.numpy()))
        # the partition is deterministic: rank 0 can work out the global order itself, so no ids travel at all
        order = np.concatenate([np.concatenate([np.arange(first[s], first[s + 1]) for s in p]) if p else np.zeros(0, np.int64) for p in parts])
        third = shard.gather_rows(torch.from_numpy(rows), None, dst=0, sizes=sizes, order=order if rank == 0 else None)
        assert (third is None) == (rank != 0) and (rank != 0 or np.array_equal(third.numpy(), out.numpy()))
        if rank == 0:
            total = int(first[-1])
            g = np.arange(total)
            want = np.stack([g * 1.5, np.sin(g), g % 7, -g, g * g], 1).astype(np.float64)
            assert out.shape == (total, 5)
            assert np.array_equal(out.numpy(), want)
            Path(tmp, "ok").write_text("ok")
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_gather_rows_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
