"""-m gpu: the multi-GPU result exchange — ONE ncclAllGather of the packed per-query
result block, issued by the library on the batch's stream.  Every rank's gathered copy must hold, for
every rank's block, exactly what the CPU oracle answers for those queries.

  * world 1 (runs on any GPU box): the collective path itself, against the oracle;
  * pb_group (one process, every visible device; skipped below 2 GPUs);
  * one process per GPU with pb_comm (skipped below 2 GPUs): what bench.py does under torchrun.
"""
import multiprocessing as mp

import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import DeviceBatch, Index, capi, score
from probly_search_b200 import distributed as D
from probly_search_b200 import workload as W

pytestmark = pytest.mark.gpu

N_DOCS, VOCAB, N_Q, K = 30_000, 1 << 13, 1203, 10      # a scaled cfg 3 (ragged: 1203 queries do not divide by 2, 4, 8)


def _n_dev():
    return capi.lib().pb_device_count()


def _corpus(device=0, removed=False):
    cfg = W.CONFIGS["cfg4" if removed else "cfg3"]
    wl = W.Workload(cfg, n_docs=N_DOCS, vocab=VOCAB)
    ix = Index(cfg.n_fields, device=device)
    wl.build_into(ix)
    for d in wl.removed_ordinals():
        ix.remove_document(int(d))
    return cfg, wl, ix


def _oracle_answers(cfg, wl, fq):
    o = orc.OracleIndex(cfg.n_fields)
    wl.build_into(o)
    for d in wl.removed_ordinals():
        o.remove_document(int(d))
    return o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, cfg.boosts, K, n_threads=8)


def _assert_equal(got, exp, lo, hi, off=0):
    """got[off + i] == exp[lo + i] for the block [lo, hi)."""
    n = hi - lo
    np.testing.assert_array_equal(got.n_results[off:off + n], exp["n_results"][lo:hi])
    np.testing.assert_array_equal(got.doc_digest[off:off + n], exp["doc_digest"][lo:hi])
    np.testing.assert_array_equal(got.score_digest[off:off + n], exp["score_digest"][lo:hi])
    np.testing.assert_array_equal(got.topk_n[off:off + n], exp["topk_n"][lo:hi])
    for i in range(n):
        m = int(got.topk_n[off + i])
        np.testing.assert_array_equal(got.topk_doc[off + i, :m], exp["topk_key"][lo + i, :m].astype(np.uint32))
        np.testing.assert_array_equal(got.topk_score[off + i, :m], exp["topk_score"][lo + i, :m])


def test_world1_gather_matches_oracle_and_local_fetch():
    cfg, wl, ix = _corpus(removed=True)
    fq = wl.queries(N_Q)
    exp = _oracle_answers(cfg, wl, fq)
    comm = D.Comm(D.Comm.new_unique_id(), 0, 1, 0)
    assert comm.nccl_version() >= 21800
    b = DeviceBatch(ix, fq, score.bm25.new(), cfg.boosts, top_k=K)
    b.set_gather(comm, N_Q + 5)                       # a slot larger than the block: the pad stays zero
    for _ in range(3):                                # re-running must not tear the block (stream-ordered memset/gather)
        b.run()
    st = b.stats()
    assert st["ms_gather"] > 0 and st["ms_gather"] < st["ms_total"]
    g = b.fetch_gathered()
    assert len(g.n_results) == N_Q + 5 and not g.n_results[N_Q:].any()
    _assert_equal(g, exp, 0, N_Q)
    _assert_equal(b.fetch(), exp, 0, N_Q)
    b.set_gather(None, 0)                             # detach: plain single-GPU run again
    b.run()
    assert b.stats()["ms_gather"] == 0
    _assert_equal(b.fetch(), exp, 0, N_Q)
    b.close(); comm.close()


@pytest.mark.skipif(_n_dev() < 2, reason="needs >= 2 GPUs")
def test_group_single_process_all_devices_match_oracle():
    cfg, wl, ix = _corpus(removed=True)
    fq = wl.queries(N_Q)
    exp = _oracle_answers(cfg, wl, fq)
    n = _n_dev()
    grp = D.Group(ix, list(range(n)))
    for _ in range(2):
        got = grp.query_batch_flat(fq, score.bm25.new(), cfg.boosts, K)
        _assert_equal(got, exp, 0, N_Q)
    rows = sum(grp.member_stats(r)["rows_scored"] for r in range(n))
    assert rows > 0
    grp.close()


def _rank_main(rank, world, uid, out):
    try:
        cfg, wl, ix = _corpus(device=rank)
        fq_all = wl.queries(N_Q)
        lo, hi, slot = D.slot_block(N_Q, rank, world)
        comm = D.Comm(uid, rank, world, rank)
        b = DeviceBatch(ix, fq_all.slice(lo, hi), score.bm25.new(), cfg.boosts, top_k=K)
        b.set_gather(comm, slot)
        for _ in range(3):
            b.run()
        g = b.fetch_gathered(N_Q)
        exp = _oracle_answers(cfg, wl, fq_all)
        _assert_equal(g, exp, 0, N_Q)                 # EVERY rank holds every rank's block, equal to the oracle
        out.put((rank, "ok", b.stats()["ms_gather"]))
        b.close(); comm.close(); ix.close()
    except Exception as e:                            # noqa: BLE001
        import traceback
        out.put((rank, "fail: " + traceback.format_exc()[-1500:], 0.0))


@pytest.mark.skipif(_n_dev() < 2, reason="needs >= 2 GPUs")
def test_one_process_per_gpu_gathered_blocks_match_oracle():
    world = min(_n_dev(), 8)
    uid = D.Comm.new_unique_id()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, uid, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, status, _ in sorted(res):
        assert status == "ok", f"rank {rank}: {status}"
