"""Data-parallel plumbing: one process per GPU, attention groups sharded over
ranks, ONE all-reduce of the flat gradient bucket per step (NCCL over NVLink; gloo in the CPU tests).

Lists are independent only at group granularity (the reference attends across the lists of a call),
so the unit of sharding is the group.  Rank r owns groups r, r+W, r+2W, ...; the reduced gradient is
the mean over all groups = the reference trained with DistributedDataParallel at per-replica batch S.
Inference and evaluation need no communication.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_groups(n_groups: int, rank: int, world: int) -> list[int]:
    """Global group indices owned by `rank` (round-robin, so every rank gets floor or ceil of n/W)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_groups, world))


def shard_lists(x: torch.Tensor, y: torch.Tensor, group_size: int, rank: int, world: int):
    """Slice a [n_groups*S, L, F] dataset into the lists of this rank's groups (group-contiguous)."""
    n_groups = x.shape[0] // group_size
    idx = shard_groups(n_groups, rank, world)
    rows = torch.cat([torch.arange(g * group_size, (g + 1) * group_size) for g in idx]) if idx else torch.empty(0, dtype=torch.long)
    rows = rows.to(x.device)
    return x.index_select(0, rows), y.index_select(0, rows)


def allreduce_mean_(bucket: torch.Tensor, groups_local: int, groups_total: int, group=None) -> torch.Tensor:
    """In-place: bucket holds the mean gradient over this rank's `groups_local` groups; afterwards it holds the
    mean over all `groups_total` groups on every rank (weights by group count, so uneven shards are exact)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        bucket.mul_(float(groups_local) / float(groups_total))
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket
