"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU path — query sharding
and the top-k gather layout (the collective the GPU run does over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from probly_search_b200 import distributed as D
from probly_search_b200 import workload as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_total, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(q_total, rank, world)
    # fake per-query results that encode (global query id, slot)
    ids = torch.arange(lo, hi, dtype=torch.int32)
    n = (ids % (k + 1)).to(torch.int32)
    docs = (ids[:, None] * 100 + torch.arange(k, dtype=torch.int32)[None, :]).to(torch.int32)
    scores = docs.to(torch.float64) * 0.5
    gn, gd, gs = D.gather_topk(n, docs, scores)
    if rank == 0:
        out.put((gn.numpy(), gd.numpy(), gs.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100_000, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [D.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gather_topk_world2_gloo():
    world, q_total, k = 2, 64, 5
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q_total, k, out)) for r in range(world)]
    for p in procs:
        p.start()
    gn, gd, gs = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids = np.arange(q_total, dtype=np.int32)
    np.testing.assert_array_equal(gn, ids % (k + 1))
    np.testing.assert_array_equal(gd, ids[:, None] * 100 + np.arange(k, dtype=np.int32)[None, :])
    np.testing.assert_array_equal(gs, gd.astype(np.float64) * 0.5)


def test_rank_query_blocks_are_disjoint_prefixes_of_one_stream():
    """bench.py gives rank r the r-th consecutive block of the config's query stream."""
    wl = W.Workload(W.CONFIGS["cfg1"], n_docs=10, vocab=256)
    whole = wl.queries(60)
    for r in range(3):
        blk = whole.slice(20 * r, 20 * (r + 1))
        assert [blk.terms_of(i) for i in range(20)] == [whole.terms_of(20 * r + i) for i in range(20)]
