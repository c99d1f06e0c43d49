"""-m gpu: differential parity of the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs — random small corpora with every awkward feature the reference handles
(empty tokens, multi-valued fields, unicode, repeated query terms, prefix expansion, boosts
<= 0, removed-but-not-vacuumed docs, vacuum, re-adding) and scaled-down BASELINE configs.
Bar: doc-id sets bit-exact; scores bit-exact (the north-star bar is 1e-9; we assert equality)."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import DeviceBatch, FlatQueries, Index, score
from probly_search_b200 import workload as W
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOK = orc.tokenizer
CALC = {orc.BM25: score.bm25.new, orc.ZERO_TO_ONE: score.zero_to_one.new}


def both(docs, n_fields):
    ix, o = Index(n_fields), orc.OracleIndex(n_fields)
    for key, fields in docs:
        ix.add_document([(lambda d, i=i: d[i]) for i in range(n_fields)], TOK, key, fields)
        o.add_document(key, fields)
    return ix, o


def compare_queries(ix, o, queries, boosts, ctx):
    for scorer in (orc.BM25, orc.ZERO_TO_ONE):
        fq = FlatQueries.from_strings(queries, TOK)
        qi, docs, scores = ix.query_full_flat(fq, CALC[scorer](), boosts)
        br = ix.query_batch_flat(fq, CALC[scorer](), boosts, top_k=5)
        for q, text in enumerate(queries):
            exp = o.query(text, scorer, boosts)
            sel = qi == q
            got = [(ix._key_of_ord(int(d)), float(s)) for d, s in zip(docs[sel], scores[sel])]
            H.assert_same_results(got, exp, ctx=f"{ctx} scorer={scorer} q={text!r} boosts={boosts}")
            # batch outputs: count, digests, top-k
            assert int(br.n_results[q]) == len(exp)
            key2ord = {ix._key_of_ord(int(d)): int(d) for d in docs[sel]}
            dd = sum(orc.doc_hash(key2ord[k]) for k, _ in exp) & (2**64 - 1)
            sd = sum(orc.score_hash(key2ord[k], s) for k, s in exp) & (2**64 - 1)
            assert int(br.doc_digest[q]) == dd
            assert int(br.score_digest[q]) == sd
            n = int(br.topk_n[q])
            assert n == min(5, len(exp))
            exp_top = sorted(((key2ord[k], s) for k, s in exp), key=lambda r: (-r[1], r[0]))[:5]
            assert [(int(br.topk_doc[q, i]), float(br.topk_score[q, i])) for i in range(n)] == exp_top


@pytest.mark.parametrize("seed", range(8))
def test_random_small_corpora(seed):
    rng = random.Random(1000 + seed)
    n_fields = rng.choice([1, 2, 2, 3, 4])
    docs = H.random_corpus(rng, rng.randint(1, 60), n_fields, multi_value=(seed % 3 == 0))
    ix, o = both(docs, n_fields)
    queries = [H.random_query(rng) for _ in range(40)] + ["", " ", "a a a", "ab abc abcd a"]
    boosts = [1.0] * n_fields
    compare_queries(ix, o, queries, boosts, f"seed={seed}")
    # boosts != 1, including zero and negative (BM25 None events, SURVEY §3.4 rule 4)
    boosts2 = [rng.choice([2.0, 0.5, 0.0, -1.0, 1.5]) for _ in range(n_fields)]
    compare_queries(ix, o, queries[:20], boosts2, f"seed={seed} boosts")
    # remove without vacuum (mask path), then vacuum, then add more
    victims = [k for k, _ in docs if rng.random() < 0.3]
    for k in victims:
        ix.remove_document(k)
        o.remove_document(k)
    compare_queries(ix, o, queries[:25], boosts, f"seed={seed} removed")
    compare_queries(ix, o, queries[:10], boosts2, f"seed={seed} removed+boosts")
    ix.vacuum()
    o.vacuum()
    compare_queries(ix, o, queries[:25], boosts, f"seed={seed} vacuumed")
    for key, fields in H.random_corpus(rng, 8, n_fields):
        ix.add_document([(lambda d, i=i: d[i]) for i in range(n_fields)], TOK, 500 + key, fields)
        o.add_document(500 + key, fields)
    compare_queries(ix, o, queries[:25], boosts, f"seed={seed} re-added")


def test_device_expansion_matches_oracle():
    rng = random.Random(7)
    docs = H.random_corpus(rng, 50, 2)
    ix, o = both(docs, 2)
    for p in ["a", "ab", "abc", "b", "x", "xy", "z", "h", "hé", "日", "日本", "nomatch", "t", "th", "the", "o"]:
        assert ix.expand_term(p) == o.expand_term(p), p


def test_live_df_kernel_matches_oracle():
    rng = random.Random(11)
    docs = H.random_corpus(rng, 80, 2)
    ix, o = both(docs, 2)
    for k, _ in docs[::3]:
        ix.remove_document(k)
        o.remove_document(k)
    df = ix.term_df_live()
    a = H.image_arrays(ix.flatten())
    removed = {k for k, _ in docs[::3]}
    for t in range(len(df)):
        term = H.image_term_string(a, t)
        live_ptrs = [k for k, _ in o.postings(term) if k not in removed]
        assert int(df[t]) == len(live_ptrs), term      # count_documents, index.rs:282-297


def _scaled(cfg_name, n_docs, vocab, n_queries, removed=False):
    cfg = W.CONFIGS[cfg_name]
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    ix, o = Index(cfg.n_fields), orc.OracleIndex(cfg.n_fields)
    wl.build_into(ix)
    wl.build_into(o)
    if removed:
        for d in wl.removed_ordinals():
            ix.remove_document(int(d))
            o.remove_document(int(d))
    fq = wl.queries(n_queries)
    scorer = orc.BM25 if cfg.scorer == "bm25" else orc.ZERO_TO_ONE
    return cfg, ix, o, fq, scorer


@pytest.mark.parametrize("cfg_name,n_docs,vocab,n_queries,removed", [
    ("cfg0", 20_000, 1 << 12, 300, False),
    ("cfg0_bench", 20_000, 1 << 12, 100, False),
    ("cfg1", 30_000, 1 << 12, 400, False),
    ("cfg2", 20_000, 1 << 12, 60, False),
    ("cfg4", 30_000, 1 << 12, 300, True),
])
def test_scaled_configs_match_oracle(cfg_name, n_docs, vocab, n_queries, removed):
    cfg, ix, o, fq, scorer = _scaled(cfg_name, n_docs, vocab, n_queries, removed)
    k = 10
    batch = DeviceBatch(ix, fq, CALC[scorer](), cfg.boosts, top_k=k)
    batch.run()
    got = batch.fetch()
    st = batch.stats()
    exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer, cfg.boosts, k)
    # keys == ordinals for the synthetic corpora, so digests and top-k compare directly
    np.testing.assert_array_equal(got.n_results, exp["n_results"])
    np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
    np.testing.assert_array_equal(got.score_digest, exp["score_digest"])     # score BITS
    np.testing.assert_array_equal(got.topk_n, exp["topk_n"])
    for q in range(fq.n_queries):
        n = int(got.topk_n[q])
        np.testing.assert_array_equal(got.topk_doc[q, :n], exp["topk_key"][q, :n].astype(np.uint32))
        np.testing.assert_array_equal(got.topk_score[q, :n], exp["topk_score"][q, :n])
    assert st["results_emitted"] == int(exp["n_results"].sum())
    # reference-equivalent pointer visits = ScoreCalculator::score calls of the oracle
    assert st["pointer_visits"] == exp["score_calls"]
    assert st["gpu_launches"] > 0
    # a second run of the same device-resident batch gives identical outputs (idempotence)
    batch.run()
    again = batch.fetch()
    np.testing.assert_array_equal(again.doc_digest, got.doc_digest)
    np.testing.assert_array_equal(again.score_digest, got.score_digest)
    np.testing.assert_array_equal(again.topk_doc, got.topk_doc)


def test_full_results_sample_matches_oracle():
    cfg, ix, o, fq, scorer = _scaled("cfg1", 30_000, 1 << 12, 40)
    qi, docs, scores = ix.query_full_flat(fq, CALC[scorer](), cfg.boosts)
    for q in range(fq.n_queries):
        exp = o.query_tokens(fq.terms_of(q), scorer, cfg.boosts)
        sel = qi == q
        H.assert_same_results([(int(d), float(s)) for d, s in zip(docs[sel], scores[sel])], exp, ctx=f"q={q}")


def test_docs_with_many_events():
    """A doc hit by more events than a warp window holds (> 32) takes the slow fold path."""
    docs = [(0, [["abc xyz b"], ["abc abc b"]]), (1, [["xyz"], ["abc b"]]), (2, [["zz"], ["b"]])]
    ix, o = both(docs, 2)
    # "abc", "xyz", "b", "zz" expand only to themselves here: events per doc = matching query terms
    queries = [" ".join(["abc"] * 40), " ".join(["abc", "b"] * 18), " ".join(["abc"] * 33 + ["b"] * 5 + ["xyz"] * 7)]
    compare_queries(ix, o, queries, [1.0, 1.0], "many-events")
    compare_queries(ix, o, queries, [0.5, -1.0], "many-events boosts")
    from probly_search_b200 import capi
    with pytest.raises(capi.ProblyError) as e:       # ZeroToOne envelope: <= 64 events per doc
        ix.query_batch_flat(FlatQueries.from_strings([" ".join(["abc"] * 70)], TOK), score.zero_to_one.new(), [1.0, 1.0], top_k=1)
    assert e.value.code == capi.PB_ERR_UNSUPPORTED


def test_empty_and_degenerate_batches():
    ix, o = both([(0, [["a b"], ["c"]]), (1, [["b"], [""]])], 2)
    fq = FlatQueries.from_strings([], TOK)
    r = ix.query_batch_flat(fq, score.bm25.new(), [1.0, 1.0], top_k=4)
    assert len(r.n_results) == 0
    r = ix.query_batch_flat(FlatQueries.from_strings(["", "zzz", "b"], TOK), score.bm25.new(), [1.0, 1.0], top_k=0)
    assert list(r.n_results) == [0, 0, 2]
    # an index with no documents at all
    empty = Index(1)
    assert empty.query("a", score.bm25.new(), TOK, [1.0]) == []


def test_error_paths_do_not_abort():
    from probly_search_b200 import capi
    ix, _ = both([(0, [["a"]])], 1)
    with pytest.raises(capi.ProblyError) as e:
        ix.query_batch_flat(FlatQueries.from_strings(["a"], TOK), score.bm25.new(), [1.0, 2.0], top_k=1)
    assert e.value.code == capi.PB_ERR_INVALID
    with pytest.raises(capi.ProblyError) as e:
        ix.query_batch_flat(FlatQueries.from_strings(["a"], TOK), score.bm25.new(), [1.0], top_k=99)
    assert e.value.code == capi.PB_ERR_UNSUPPORTED
    with pytest.raises(capi.ProblyError):
        ix.query_batch_flat(FlatQueries.from_strings(["a"], TOK), score.bm25.new(), [float("nan")], top_k=1)
