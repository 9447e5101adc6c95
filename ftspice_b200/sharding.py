"""Multi-GPU plumbing: one process per GPU, instances sharded contiguously, no data-path collective.

Instances (and .dc chains) are independent, so the batch is cut into contiguous blocks of ceil(B/G) instances
(SURVEY.md §8e).  The only exchange is the final gather of results (operating points, final states, counts,
status) — a plain all_gather over NCCL (NVLink) on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, first + count) of the global instance indices owned by `rank`."""
    per = (total + world - 1) // world
    first = min(rank * per, total)
    return first, max(0, min(per, total - first))


def gather_rows(local: torch.Tensor, total: int) -> torch.Tensor:
    """All ranks' row blocks (contiguous shards of a [total, ...] array) -> the full array on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = (total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad) if local.is_cuda else dist.all_gather(list(out.chunk(world)), pad)
    return out[:total]


def max_over_ranks(values: List[float], device=None) -> List[float]:
    """Device times are reported as the max over ranks."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
