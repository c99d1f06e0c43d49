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


@pytest.mark.parametrize("row_dead", ["1", "0"])
@pytest.mark.parametrize("cfg_name,removed_cfg", [("cfg4", "cfg4"), ("cfg2", "cfg4"), ("cfg1", "cfg4")])
def test_removed_docs_by_row_bits_and_by_bitmap_probe(monkeypatch, row_dead, cfg_name, removed_cfg):
    """Removed-but-not-vacuumed docs (query.rs:65): the scoring loop reads their bits per posting ROW with the tile
    (IndexView::row_dead, default) or probes the doc bitmap (PB_ROW_DEAD=0; edge tiles and the wide layout always do).
    Both must give the oracle's answers; a second set_live_state (more removals) must rebuild the row bits."""
    monkeypatch.setenv("PB_ROW_DEAD", row_dead)
    cfg = W.CONFIGS[cfg_name]
    n_docs, vocab = 60_000, 1 << 11
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    ix, o = Index(cfg.n_fields), orc.OracleIndex(cfg.n_fields)
    wl.build_into(ix)
    wl.build_into(o)
    fq = wl.queries(120)
    scorer = orc.BM25 if cfg.scorer == "bm25" else orc.ZERO_TO_ONE
    gone = W.Workload(W.CONFIGS[removed_cfg], n_docs=n_docs, vocab=vocab).removed_ordinals()
    k = 10
    for part in (gone[: len(gone) // 2], gone[len(gone) // 2:]):
        for d in part:
            ix.remove_document(int(d))
            o.remove_document(int(d))
        got = ix.query_batch_flat(fq, CALC[scorer](), cfg.boosts, k)
        exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer, cfg.boosts, k)
        np.testing.assert_array_equal(got.n_results, exp["n_results"])
        np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
        np.testing.assert_array_equal(got.score_digest, exp["score_digest"])
        np.testing.assert_array_equal(got.topk_n, exp["topk_n"])
        for q in range(fq.n_queries):
            n = int(got.topk_n[q])
            np.testing.assert_array_equal(got.topk_doc[q, :n], exp["topk_key"][q, :n].astype(np.uint32))
            np.testing.assert_array_equal(got.topk_score[q, :n], exp["topk_score"][q, :n])
        assert ix.last_stats()["pointer_visits"] == exp["score_calls"]


@pytest.mark.parametrize("boosts", [[2.0, 0.5], [4.0, 0.25], [-2.0, 1.5], [0.5, 3.0], [1.0, 0.125]])
def test_power_of_two_boosts_folded_into_the_table(monkeypatch, boosts):
    """A boost of +-2^k is folded into the table of saturated tf (ScoreParams::tab_scale): exact, so the scores must be
    the oracle's bit for bit, folded (default) or not (PB_FOLD_BOOST=0), for mixed folded / unfolded fields, and when
    the same staged batch's boosts change between runs."""
    cfg, ix, o, fq, scorer = _scaled("cfg1", 30_000, 1 << 11, 150)
    exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, boosts, 10)
    for fold in ("1", "0", "1"):
        monkeypatch.setenv("PB_FOLD_BOOST", fold)
        got = ix.query_batch_flat(fq, score.bm25.new(), boosts, 10)
        np.testing.assert_array_equal(got.n_results, exp["n_results"])
        np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
        np.testing.assert_array_equal(got.score_digest, exp["score_digest"])
        for q in range(fq.n_queries):
            n = int(got.topk_n[q])
            np.testing.assert_array_equal(got.topk_doc[q, :n], exp["topk_key"][q, :n].astype(np.uint32))
            np.testing.assert_array_equal(got.topk_score[q, :n], exp["topk_score"][q, :n])
        # unit boosts next through the same scratch batch: the folded table must not survive
        one = ix.query_batch_flat(fq, score.bm25.new(), [1.0, 1.0], 10)
        e1 = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, [1.0, 1.0], 10)
        np.testing.assert_array_equal(one.score_digest, e1["score_digest"])


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
    # no cap on the events one doc may receive (round 1 refused > 64 in a ZeroToOne query)
    compare_queries(ix, o, [" ".join(["abc"] * 70), " ".join(["abc", "b", "xyz"] * 45)], [1.0, 1.0], "many-events > 64")


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


# ---- device posting layouts: results must not depend on the layout ------------
@pytest.mark.parametrize("layout", ["wide", "auto"])
@pytest.mark.parametrize("seed", [3, 6])
def test_random_corpora_both_layouts(monkeypatch, layout, seed):
    monkeypatch.setenv("PB_POSTING_LAYOUT", layout)
    rng = random.Random(7000 + seed)
    n_fields = [1, 2, 3, 4][seed % 4]
    docs = H.random_corpus(rng, rng.randint(20, 60), n_fields, multi_value=True)
    ix, o = both(docs, n_fields)
    queries = [H.random_query(rng) for _ in range(30)] + ["a", "ab abc abcd a"]
    compare_queries(ix, o, queries, [1.0] * n_fields, f"layout={layout} seed={seed}")
    compare_queries(ix, o, queries[:12], [rng.choice([2.0, 0.5, -1.0]) for _ in range(n_fields)], f"layout={layout} boosts")
    assert ix.device_layout()["narrow"] == (layout == "auto")
    for k, _ in docs[::3]:
        ix.remove_document(k)
        o.remove_document(k)
    compare_queries(ix, o, queries[:15], [1.0] * n_fields, f"layout={layout} removed")


@pytest.mark.parametrize("layout", ["wide", "auto"])
@pytest.mark.parametrize("cfg_name,n_docs,vocab,n_queries,removed", [
    ("cfg1", 30_000, 1 << 12, 300, False),
    ("cfg2", 20_000, 1 << 12, 50, False),
    ("cfg4", 30_000, 1 << 12, 200, True),
])
def test_scaled_configs_both_layouts(monkeypatch, layout, cfg_name, n_docs, vocab, n_queries, removed):
    monkeypatch.setenv("PB_POSTING_LAYOUT", layout)
    cfg, ix, o, fq, scorer = _scaled(cfg_name, n_docs, vocab, n_queries, removed)
    batch = DeviceBatch(ix, fq, CALC[scorer](), cfg.boosts, top_k=10)
    batch.run()
    got = batch.fetch()
    assert ix.device_layout()["narrow"] == (layout == "auto")
    assert ix.device_layout()["bytes_per_row"] == (4 + 2 * cfg.n_fields if layout == "auto" else 4 + 8 * cfg.n_fields)
    exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer, cfg.boosts, 10)
    np.testing.assert_array_equal(got.n_results, exp["n_results"])
    np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
    np.testing.assert_array_equal(got.score_digest, exp["score_digest"])
    for q in range(fq.n_queries):
        n = int(got.topk_n[q])
        np.testing.assert_array_equal(got.topk_doc[q, :n], exp["topk_key"][q, :n].astype(np.uint32))
        np.testing.assert_array_equal(got.topk_score[q, :n], exp["topk_score"][q, :n])


@pytest.mark.parametrize("seed", [1, 4, 5])
def test_rank_directory_for_every_list(monkeypatch, seed):
    """The marking pass finds secondary docs in a dense primary list through the rank directory; forcing a
    directory for every list (PB_DIR_MIN_ROWS=1) runs that path on small random corpora."""
    monkeypatch.setenv("PB_DIR_MIN_ROWS", "1")
    rng = random.Random(8000 + seed)
    n_fields = rng.choice([1, 2, 3])
    docs = H.random_corpus(rng, rng.randint(30, 80), n_fields, multi_value=(seed % 2 == 0))
    ix, o = both(docs, n_fields)
    queries = [H.random_query(rng) for _ in range(40)] + ["a", "ab", "a b", "ab abc abcd a"]
    compare_queries(ix, o, queries, [1.0] * n_fields, f"dir seed={seed}")
    for k, _ in docs[::4]:
        ix.remove_document(k)
        o.remove_document(k)
    compare_queries(ix, o, queries[:20], [1.0] * n_fields, f"dir seed={seed} removed")


def test_long_fields_fall_back_to_wide_layout():
    """A field of 300 tokens / a tf of 300 does not fit a u16 (tf, fl) code: the device keeps u32 columns
    (and the BM25 table no longer covers every (tf, fl): the exact division path runs)."""
    rng = random.Random(99)
    words = ["alpha", "beta", "gamma", "delta", "al", "alp", "be"]
    docs = []
    for k in range(40):
        n0 = rng.choice([3, 10, 300, 700])
        f0 = " ".join(rng.choice(words) for _ in range(n0))
        f1 = " ".join(["alpha"] * rng.choice([1, 2, 300])) + " " + " ".join(rng.choice(words) for _ in range(5))
        docs.append((k, [[f0], [f1]]))
    ix, o = both(docs, 2)
    compare_queries(ix, o, ["alpha", "al", "be gamma", "a", "alpha alpha beta", "delta alp"], [1.0, 1.0], "long fields")
    compare_queries(ix, o, ["alpha", "al be"], [0.5, 2.0], "long fields boosts")
    assert ix.device_layout()["narrow"] is False


def test_index_served_from_an_image_file(tmp_path):
    """save the flattened image, load it in a builder-less Index, same results bit for bit
    (pre-vacuum removed docs included: the live state travels in the image)."""
    cfg, ix, o, fq, scorer = _scaled("cfg4", 20_000, 1 << 12, 150, removed=True)
    p = str(tmp_path / "cfg4.pbimg")
    ix.save_image(p)
    ld = Index.load_image(p)
    a = ix.query_batch_flat(fq, CALC[scorer](), cfg.boosts, top_k=10)
    b = ld.query_batch_flat(fq, CALC[scorer](), cfg.boosts, top_k=10)
    exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer, cfg.boosts, 10)
    for got in (a, b):
        np.testing.assert_array_equal(got.n_results, exp["n_results"])
        np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
        np.testing.assert_array_equal(got.score_digest, exp["score_digest"])
    np.testing.assert_array_equal(a.topk_doc, b.topk_doc)
    np.testing.assert_array_equal(a.topk_score, b.topk_score)
    assert ld.expand_term("a") == ix.expand_term("a")
    ld.close()


def test_one_call_queries_from_several_host_threads():
    """pb_query_batch from 4 host threads at once on ONE index (each call borrows an internal batch, own stream):
    every answer must be the oracle's, whatever the interleaving; a removal in between excludes running queries."""
    import threading
    cfg, ix, o, fq, scorer = _scaled("cfg1", 40_000, 1 << 11, 96)
    k = 5
    exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer, cfg.boosts, k)
    ix.query_batch_flat(fq.slice(0, 1), CALC[scorer](), cfg.boosts, k)         # image resident before the threads start
    errors = []

    def worker(t):
        try:
            for rep in range(3):
                for q in range(t, fq.n_queries, 4):
                    n = 1 if (q + rep) % 3 else min(7, fq.n_queries - q)     # single queries and small batches mixed
                    r = ix.query_batch_flat(fq.slice(q, q + n), CALC[scorer](), cfg.boosts, k)
                    for j in range(n):
                        assert int(r.n_results[j]) == int(exp["n_results"][q + j])
                        assert int(r.doc_digest[j]) == int(exp["doc_digest"][q + j])
                        assert int(r.score_digest[j]) == int(exp["score_digest"][q + j])
                        m = int(r.topk_n[j])
                        assert m == int(exp["topk_n"][q + j])
                        assert r.topk_doc[j, :m].tolist() == exp["topk_key"][q + j, :m].astype(np.uint32).tolist()
        except Exception as e:          # surfaced in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert ix.last_stats()["n_queries"] >= 1


def test_live_state_change_while_other_threads_query():
    """pb_index_set_live_state takes the index exclusively: a one-call query running in another host thread sees the
    state before or after it, never a mixture (idf / avg / removed mask / BM25 table of different states)."""
    import threading
    cfg = W.CONFIGS["cfg1"]
    n_docs, vocab = 40_000, 1 << 11
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    ix, ix2, o = Index(cfg.n_fields), Index(cfg.n_fields), orc.OracleIndex(cfg.n_fields)
    for x in (ix, ix2, o):
        wl.build_into(x)
    fq = wl.queries(64)
    k = 5
    before = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, cfg.boosts, k)
    gone = W.Workload(W.CONFIGS["cfg4"], n_docs=n_docs, vocab=vocab).removed_ordinals()
    for d in gone:
        ix2.remove_document(int(d))
        o.remove_document(int(d))
    after = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, cfg.boosts, k)
    ords, n_live, avg = ix2.live_state()
    ix.query_batch_flat(fq.slice(0, 1), score.bm25.new(), cfg.boosts, k)
    errors, seen = [], {"before": 0, "after": 0}
    stop = threading.Event()

    def worker(t):
        try:
            q = t
            while not stop.is_set():
                r = ix.query_batch_flat(fq.slice(q, q + 1), score.bm25.new(), cfg.boosts, k)
                got = (int(r.n_results[0]), int(r.doc_digest[0]), int(r.score_digest[0]))
                b = (int(before["n_results"][q]), int(before["doc_digest"][q]), int(before["score_digest"][q]))
                a = (int(after["n_results"][q]), int(after["doc_digest"][q]), int(after["score_digest"][q]))
                assert got == b or got == a, (q, got, b, a)
                seen["before" if got == b else "after"] += 1
                q = (q + 3) % fq.n_queries
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    import time as _t
    _t.sleep(0.05)
    ix.set_live_state(ords, n_live, avg)          # C-level call only: the host mirror of `ix` is not touched
    _t.sleep(0.05)
    stop.set()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert seen["after"] > 0
