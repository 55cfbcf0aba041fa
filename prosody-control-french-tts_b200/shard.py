"""Multi-GPU plumbing: one process per GPU, units sharded by segment, one gather of the per-syntagme rows.

Every (file, t0, t1) unit is independent, so the path needs no data-path collective: segments are
dealt to ranks by length-balanced bucketing (longest first onto the least-loaded rank; a segment's whole-file and
syntagme units stay together so its PCM is uploaded once), each rank measures its shard on its own GPU, and the
fixed-width result rows are gathered on rank 0 — the only exchange.  The cross-unit steps (sliding-window baselines,
EMA over all rows of a voice) are host-side and run on rank 0 after the gather.

`torch.distributed` is plumbing only: NCCL over NVLink for CUDA tensors, gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def partition_by_cost(costs: Sequence[float], world: int) -> list[list[int]]:
    """Greedy longest-first bucketing. costs[i] ~ work of segment i (e.g. its sample or frame count).
    Returns, per rank, the segment indices it owns in ascending order (segment order matters to the EMA later)."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(costs), np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return [sorted(np.nonzero(owner == r)[0].tolist()) for r in range(world)]


def row_counts(n_local: int, device) -> list[int]:
    """Row count of every rank (one all_gather).  A caller whose shards do not change from step to step exchanges them
    once and passes them to gather_rows, which then needs no host synchronisation on the sending ranks."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, mine)
    return [int(s.item()) for s in sizes]


def gather_rows(rows, index, dst: int = 0, sizes: list[int] | None = None):
    """Gather [n_r, k] float64 rows and their global row ids [n_r] from every rank onto `dst`; returns the rows in
    global order on dst (None elsewhere).  Works for gloo (CPU tensors) and NCCL (CUDA tensors)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = rows.device
    k = rows.shape[1]
    if sizes is None:
        sizes = row_counts(rows.shape[0], dev)
    mx = max(sizes + [1])
    pad = torch.zeros(mx, k + 1, dtype=torch.float64, device=dev)
    pad[:rows.shape[0], :k] = rows
    pad[:rows.shape[0], k] = index.to(torch.float64)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    allr = torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)
    ids = allr[:, k].to(torch.int64)
    # rows come sorted per rank; only an interleaved partition needs the global sort (done where the data lives: on the GPU
    # for NCCL, where a 5e5-row argsort is microseconds instead of the tens of milliseconds of a single host thread)
    if ids.numel() > 1 and not bool((ids[1:] >= ids[:-1]).all()):
        allr = allr[torch.argsort(ids, stable=True)]
    return allr[:, :k].cpu()
