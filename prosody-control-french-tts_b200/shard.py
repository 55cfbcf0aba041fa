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


def row_counts(n_local: int, device, group=None) -> list[int]:
    """Row count of every rank (ONE all_gather_into_tensor).  A caller whose shards do not change from step to step exchanges them
    once and passes them to gather_rows, which then needs no size exchange and no host synchronisation on the sending ranks."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    return [int(v) for v in sizes.tolist()]


def gather_rows(rows, index=None, dst: int = 0, sizes: list[int] | None = None, order=None, perm=None, out=None, group=None):
    """Ragged gather of [n_r, k] float64 rows from every rank onto `dst`: every rank sends exactly its rows (point-to-point, true
    counts — nothing is padded to the largest shard), `dst` receives them rank by rank.  Global order: either `order`, the global
    row id of every gathered row in rank-major order, which `dst` can compute itself when the partition is deterministic (no ids
    travel at all), or `index`, this rank's int64 ids, which then travel as their own int64 message; or `perm`, a precomputed
    index tensor on the rows' device (rank-major position of every global row: the order worked out once for shards that do not
    change).  Returns the rows in global order on dst (None elsewhere): a CPU tensor, or — when `out` (a pinned CPU tensor of the
    full size) is given — `out`, filled by an asynchronous copy the caller synchronises.  Works for gloo and NCCL; `group` selects
    the process group (the rows of the step are computed on the host: a gloo group moves them without touching the GPUs, whose
    SMs are held by persistent kernels that a NCCL send / receive kernel would have to wait for)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = rows.device
    k = rows.shape[1]
    if sizes is None:
        sizes = row_counts(rows.shape[0], dev, group)
    rows = rows.contiguous()
    send_ids = index is not None and order is None
    if rank != dst:
        ops = []
        if rows.shape[0]:
            ops.append(dist.P2POp(dist.isend, rows, dst, group))
            if send_ids:
                ops.append(dist.P2POp(dist.isend, index.to(torch.int64).contiguous(), dst, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return None
    parts = [rows if r == dst else torch.empty(sizes[r], k, dtype=torch.float64, device=dev) for r in range(world)]
    idp = [(index.to(torch.int64) if r == dst else torch.empty(sizes[r], dtype=torch.int64, device=dev)) for r in range(world)] if send_ids else None
    ops = []
    for r in range(world):
        if r != dst and sizes[r]:
            ops.append(dist.P2POp(dist.irecv, parts[r], r, group))
            if send_ids:
                ops.append(dist.P2POp(dist.irecv, idp[r], r, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    allr = torch.cat(parts, 0)
    ids = None
    if order is not None:
        ids = torch.as_tensor(order, dtype=torch.int64, device=dev)
    elif send_ids:
        ids = torch.cat(idp, 0)
    # rows come sorted per rank; only an interleaved partition needs the global sort (done where the data lives: on the GPU
    # for NCCL, where a 5e5-row argsort is microseconds instead of the tens of milliseconds of a single host thread)
    if perm is not None:
        allr = allr.index_select(0, perm)
    elif ids is not None and ids.numel() > 1 and not bool((ids[1:] >= ids[:-1]).all()):
        allr = allr[torch.argsort(ids, stable=True)]
    if out is not None:
        out.copy_(allr, non_blocking=True)
        return out
    return allr.cpu()
