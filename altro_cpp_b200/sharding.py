"""Batch sharding across GPUs of one node.

Instances are fully independent, so the path shards trivially over the batch axis: contiguous
slices, one process per GPU, NO collective on the data path.  torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) is used only to scatter the per-instance inputs
from rank 0 and to gather the per-instance results back — plumbing, outside the timed region.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) slices, sizes differing by at most one (ragged batches allowed)."""
    base, rem = divmod(total, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def scatter_rows(full: Optional[torch.Tensor], total: int, row_shape, dtype, device, src: int = 0) -> torch.Tensor:
    """Rank `src` holds `full` [total, *row_shape]; every rank returns its own slice."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(total, world)
    width = max(hi - lo for lo, hi in bounds)
    recv = torch.empty((width, *row_shape), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for lo, hi in bounds:
            c = torch.zeros((width, *row_shape), dtype=dtype, device=device)
            c[: hi - lo] = full[lo:hi].to(device)
            chunks.append(c)
    dist.scatter(recv, chunks, src=src)
    lo, hi = bounds[rank]
    return recv[: hi - lo].contiguous()


def gather_rows(local: torch.Tensor, total: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Inverse of scatter_rows: rank `dst` gets [total, *row_shape], the others None."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(total, world)
    width = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((width, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, bounds)], dim=0)
