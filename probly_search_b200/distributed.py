"""Multi-GPU plumbing for the query path: queries are independent units, so a
batch is sharded by query over ranks that each hold a replica of the index image; the ONLY
collective is the gather of the fixed-size per-query top-k blocks (torch.distributed: NCCL over
NVLink on GPUs, gloo in the CPU tests).  Nothing here touches the scoring path."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: [lo, hi) with sizes differing by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DeviceArray:
    """Wraps a raw device pointer (e.g. from pb_batch_device_results) for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def gather_topk(local_n: torch.Tensor, local_docs: torch.Tensor, local_scores: torch.Tensor, group=None):
    """All-gathers per-query top-k blocks of equally sized shards.
    local_n [q], local_docs [q, k], local_scores [q, k]  ->  ([W*q], [W*q, k], [W*q, k]) in rank order."""
    world = dist.get_world_size(group)
    q, k = local_docs.shape
    out_n = torch.empty(world * q, dtype=local_n.dtype, device=local_n.device)
    out_d = torch.empty(world * q, k, dtype=local_docs.dtype, device=local_docs.device)
    out_s = torch.empty(world * q, k, dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out_n, local_n.contiguous(), group=group)
    dist.all_gather_into_tensor(out_d.view(-1), local_docs.contiguous().view(-1), group=group)
    dist.all_gather_into_tensor(out_s.view(-1), local_scores.contiguous().view(-1), group=group)
    return out_n, out_d, out_s
