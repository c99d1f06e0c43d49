"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU path — the query
partition and the layout of the gathered result blocks (on GPUs the library itself does this
exchange with ONE ncclAllGather of the packed block, see tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from probly_search_b200 import distributed as D
from probly_search_b200 import workload as W
from probly_search_b200.index import BatchResults


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_block(lo, hi, k):
    """Per-query results that encode the GLOBAL query id."""
    ids = np.arange(lo, hi, dtype=np.uint64)
    r = BatchResults(hi - lo, k)
    r.n_results[:] = ids * 3 + 1
    r.doc_digest[:] = ids * np.uint64(0x9E3779B97F4A7C15)
    r.score_digest[:] = ids ^ np.uint64(0xABCDEF)
    r.topk_n[:] = (ids % (k + 1)).astype(np.uint32)
    r.topk_doc[:] = (ids[:, None] * 100 + np.arange(k, dtype=np.uint64)[None, :]).astype(np.uint32)
    r.topk_score[:] = r.topk_doc.astype(np.float64) * 0.5
    return r


def _worker(rank, world, port, q_total, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi, slot = D.slot_block(q_total, rank, world)
    g = D.gather_blocks_host(_fake_block(lo, hi, k), slot)
    if rank == 0:
        out.put({n: getattr(g, n) for n in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score")})
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


def test_slot_block_is_the_librarys_partition():
    """pb_group_query_batch cuts a batch into blocks of ceil(n / world): global query g = rank * slot + i."""
    for n in (0, 1, 7, 63, 100_000, 1_000_003):
        for w in (1, 2, 3, 8):
            blocks = [D.slot_block(n, r, w) for r in range(w)]
            slot = blocks[0][2]
            assert slot == -(-n // w) and slot * w >= n
            covered = [g for lo, hi, _ in blocks for g in range(lo, hi)] if n <= 100 else None
            if covered is not None:
                assert covered == list(range(n))
            for r, (lo, hi, s) in enumerate(blocks):
                assert s == slot and lo == min(n, r * slot) and hi - lo <= slot


def test_gathered_blocks_world2_gloo_ragged():
    """63 queries over 2 ranks: blocks of 32 and 31; the gathered arrays are [world * slot] with rank r's
    query i at r * slot + i, so the first n_total entries are the batch in the caller's order."""
    world, q_total, k = 2, 63, 5
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q_total, k, out)) for r in range(world)]
    for p in procs:
        p.start()
    g = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    exp = _fake_block(0, q_total, k)
    for name in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score"):
        np.testing.assert_array_equal(g[name][:q_total], getattr(exp, name))
        assert not g[name][q_total:].any()           # the pad of the short last block stays zero


def test_rank_query_blocks_are_disjoint_prefixes_of_one_stream():
    """bench.py gives rank r the r-th consecutive block of the config's query stream."""
    wl = W.Workload(W.CONFIGS["cfg1"], n_docs=10, vocab=256)
    whole = wl.queries(60)
    for r in range(3):
        blk = whole.slice(20 * r, 20 * (r + 1))
        assert [blk.terms_of(i) for i in range(20)] == [whole.terms_of(20 * r + i) for i in range(20)]
