"""Multi-GPU plumbing for the query path: queries are independent units, so a
batch is sharded by query over ranks that each hold a replica of the index image; the ONLY
collective is one ncclAllGather of the packed per-query result blocks, issued INSIDE the library on
the batch's own stream (include/probly_b200.h "Multi-GPU").  This module is the host-side glue:

  * `Comm`   — one rank of the library's NCCL communicator (pb_comm_*); in a torchrun job the
               128-byte NCCL id travels from rank 0 over the existing torch.distributed group;
  * `Group`  — single-process form (pb_group_*): replicas on several devices of one process;
  * `shard_range` / `gather_blocks_host` — the block arithmetic, also exercised by the gloo tests.

Nothing here touches the scoring path."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import capi
from .index import BatchResults, FlatQueries, Index


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: [lo, hi) with sizes differing by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slot_block(n_items: int, rank: int, world: int) -> Tuple[int, int, int]:
    """The library's own partition (pb_group_query_batch): blocks of slot = ceil(n / world) queries;
    returns (lo, hi, slot).  Global query g lives at rank g // slot, local index g % slot."""
    slot = (n_items + world - 1) // world if n_items else 0
    lo = min(n_items, slot * rank)
    hi = min(n_items, slot * (rank + 1))
    return lo, hi, slot


class Comm:
    """One rank of an NCCL communicator owned by the product library."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        self._L = capi.lib()
        if len(unique_id) != capi.PB_COMM_ID_BYTES:
            raise ValueError("unique_id must be PB_COMM_ID_BYTES long")
        buf = (C.c_uint8 * capi.PB_COMM_ID_BYTES).from_buffer_copy(unique_id)
        h = C.c_void_p()
        capi.check(self._L.pb_comm_create(buf, rank, world, device, C.byref(h)))
        self._h = h
        self.rank, self.world, self.device = rank, world, device

    @staticmethod
    def new_unique_id() -> bytes:
        buf = (C.c_uint8 * capi.PB_COMM_ID_BYTES)()
        capi.check(capi.lib().pb_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch(cls, device: int, group=None) -> "Comm":
        """Inside a torch.distributed job: rank 0 makes the id, a broadcast carries it."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        backend = dist.get_backend(group)
        dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
        t = torch.zeros(capi.PB_COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.new_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0, group=group)
        return cls(bytes(t.cpu().numpy().tobytes()), rank, world, device)

    def nccl_version(self) -> int:
        v = C.c_int(0)
        capi.check(self._L.pb_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.pb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """Single-process multi-GPU form: replicas of one image on `devices` (pb_group_create)."""

    def __init__(self, index: Index, devices: Sequence[int]):
        self._L = capi.lib()
        self.index = index
        im = index.flatten()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        capi.check(self._L.pb_group_create(C.byref(im), devs, len(devices), C.byref(h)))
        self._h = h
        self.devices = list(devices)

    def query_batch_flat(self, fq: FlatQueries, score_calculator, fields_boost: Sequence[float], top_k: int = 10) -> BatchResults:
        d, _keep = self.index._desc(fq, score_calculator, fields_boost, top_k)
        res = BatchResults(fq.n_queries, top_k)
        rs = res.c_struct()
        capi.check(self._L.pb_group_query_batch(self._h, C.byref(d), C.byref(rs)))
        return res

    def member_stats(self, member: int) -> dict:
        s = capi.BatchStats()
        capi.check(self._L.pb_group_member_stats(self._h, member, C.byref(s)))
        return s.as_dict()

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.pb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_blocks_host(local: BatchResults, slot: int, group=None) -> BatchResults:
    """The same exchange over a torch.distributed group with HOST tensors (gloo): every rank
    contributes a block padded to `slot` queries; returns the world * slot gathered results.  This is
    the CPU-testable mirror of pb_batch_fetch_gathered's layout — the GPU path never calls it."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    k = max(local.k, 1)
    n = len(local.n_results)
    if n > slot:
        raise ValueError("block larger than the slot")
    out = BatchResults(world * slot, local.k)
    for name in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score"):
        a = getattr(local, name)
        pad = np.zeros((slot,) + a.shape[1:], dtype=a.dtype)
        pad[:n] = a
        # gloo has no unsigned 32/64-bit types: ship the raw bytes
        t = torch.from_numpy(pad.view(np.uint8).reshape(-1).copy())
        g = torch.empty(world * t.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(g, t, group=group)
        getattr(out, name)[...] = g.numpy().view(a.dtype).reshape((world * slot,) + a.shape[1:])
    return out
