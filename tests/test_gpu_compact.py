"""-m gpu: the compact copy of the tiles (u16 doc offsets; IndexView::cpost in csrc/kernels.cuh) that the
single-list launch can stream instead of the u32 doc column.  PB_POSTING_COMPACT=1 builds it (off by default: measured
slower on B200 although it reads fewer bytes) and
PB_COMPACT_MIN_TILES=1 lets every list with one interior tile use it; the results must equal the oracle's, bit for bit,
and those of the same index without the copy.  The stats say how many rows really streamed from it."""
import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import DeviceBatch, Index, score
from probly_search_b200 import workload as W
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _build(cfg, n_docs, vocab, removed=False):
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    ix, o = Index(cfg.n_fields), orc.OracleIndex(cfg.n_fields)
    wl.build_into(ix)
    wl.build_into(o)
    if removed:
        for d in W.Workload(W.CONFIGS["cfg4"], n_docs=n_docs, vocab=vocab).removed_ordinals():
            ix.remove_document(int(d))
            o.remove_document(int(d))
    return wl, ix, o


def _check(got, exp, nq):
    np.testing.assert_array_equal(got.n_results, exp["n_results"])
    np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
    np.testing.assert_array_equal(got.score_digest, exp["score_digest"])
    np.testing.assert_array_equal(got.topk_n, exp["topk_n"])
    for q in range(nq):
        n = int(got.topk_n[q])
        np.testing.assert_array_equal(got.topk_doc[q, :n], exp["topk_key"][q, :n].astype(np.uint32))
        np.testing.assert_array_equal(got.topk_score[q, :n], exp["topk_score"][q, :n])


@pytest.mark.parametrize("name,scorer,removed", [("cfg1", "bm25", False), ("cfg4", "bm25", True), ("cfg1", "zero_to_one", False),
                                                 ("cfg1", "zero_to_one", True)])
def test_single_list_stream_from_compact_tiles(monkeypatch, name, scorer, removed):
    monkeypatch.setenv("PB_POSTING_COMPACT", "1")
    monkeypatch.setenv("PB_COMPACT_MIN_TILES", "1")
    cfg = W.CONFIGS[name]
    # 150 k docs: sparse lists have tiles spanning more than 2^16 docs (not compact), dense ones fit
    wl, ix, o = _build(cfg, 150_000, 1 << 12, removed)
    fq = wl.queries(300)
    calc = score.bm25.new() if scorer == "bm25" else score.zero_to_one.new()
    oid = orc.BM25 if scorer == "bm25" else orc.ZERO_TO_ONE
    k = 10
    batch = DeviceBatch(ix, fq, calc, cfg.boosts, top_k=k)
    for _ in range(2):
        batch.run()
        got, st = batch.fetch(), batch.stats()
        exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, oid, cfg.boosts, k, n_threads=8)
        _check(got, exp, fq.n_queries)
        assert st["pointer_visits"] == exp["score_calls"]
        assert 0 < st["rows_streamed_compact"] <= st["rows_streamed_direct"]
    batch.close()
    # the same index without the copy: identical answers, nothing streamed from it
    monkeypatch.setenv("PB_POSTING_COMPACT", "0")
    ix2 = Index(cfg.n_fields)
    wl.build_into(ix2)
    if removed:
        for d in W.Workload(W.CONFIGS["cfg4"], n_docs=150_000, vocab=1 << 12).removed_ordinals():
            ix2.remove_document(int(d))
    b2 = DeviceBatch(ix2, fq, calc, cfg.boosts, top_k=k)
    b2.run()
    got2, st2 = b2.fetch(), b2.stats()
    assert st2["rows_streamed_compact"] == 0
    for f in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score"):
        np.testing.assert_array_equal(getattr(got, f), getattr(got2, f))
    b2.close()


def test_compact_tiles_with_non_unit_boosts_and_full_results(monkeypatch):
    """Boosts [2, 0.5] take the SIMPLE = 0 instantiation; full result sets (capture) stay on the u32 tiles."""
    monkeypatch.setenv("PB_POSTING_COMPACT", "1")
    monkeypatch.setenv("PB_COMPACT_MIN_TILES", "1")
    cfg = W.CONFIGS["cfg4"]
    wl, ix, o = _build(cfg, 30_000, 1 << 10)
    fq = wl.queries(40)
    exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, cfg.boosts, 10, n_threads=8)
    got = ix.query_batch_flat(fq, score.bm25.new(), cfg.boosts, 10)
    _check(got, exp, fq.n_queries)
    assert ix.last_stats()["rows_streamed_compact"] > 0
    qi, docs, scores = ix.query_full_flat(fq, score.bm25.new(), cfg.boosts)
    for q in range(fq.n_queries):
        e = o.query_tokens(fq.terms_of(q), orc.BM25, cfg.boosts)
        sel = qi == q
        H.assert_same_results([(int(d), float(s)) for d, s in zip(docs[sel], scores[sel])], e, ctx=f"q={q}")


def test_the_copy_is_not_built_by_default():
    cfg = W.CONFIGS["cfg1"]
    wl = W.Workload(cfg, n_docs=20_000, vocab=1 << 10)
    ix = Index(cfg.n_fields)
    wl.build_into(ix)
    fq = wl.queries(20)
    ix.query_batch_flat(fq, score.bm25.new(), cfg.boosts, 5)
    assert ix.last_stats()["rows_streamed_compact"] == 0
