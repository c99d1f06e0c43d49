"""-m gpu: the dense union path (csrc/union_kernels.cuh) — ZeroToOne for union-heavy queries — against
the CPU oracle, bit for bit.  PB_UNION_MIN_DIV=0 sends EVERY eligible multi-list ZeroToOne query
through it (small corpora included), PB_UNION=0 disables it, so both routes stay covered; the stats
say which route ran."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import DeviceBatch, FlatQueries, Index, score
from probly_search_b200 import workload as W
from tests import helpers as H
from tests.test_gpu_parity import TOK, _scaled, both, compare_queries

pytestmark = pytest.mark.gpu


def _route(monkeypatch, route):
    if route == "union_all":
        monkeypatch.setenv("PB_UNION_MIN_DIV", "0")
    elif route == "union_cta":                       # the CTA-per-2048-doc-shard kernel instead of the warp-per-512-doc-shard one
        monkeypatch.setenv("PB_UNION_MIN_DIV", "0")
        monkeypatch.setenv("PB_UNION_KERNEL", "cta")
    elif route == "union_off":
        monkeypatch.setenv("PB_UNION", "0")


@pytest.mark.parametrize("route", ["union_all", "union_cta", "union_off"])
@pytest.mark.parametrize("seed", range(6))
def test_random_corpora_through_both_routes(monkeypatch, route, seed):
    _route(monkeypatch, route)
    rng = random.Random(4200 + seed)
    n_fields = [1, 2, 2, 3, 4, 2][seed]
    docs = H.random_corpus(rng, rng.randint(20, 90), n_fields, multi_value=(seed % 2 == 1))
    ix, o = both(docs, n_fields)
    queries = [H.random_query(rng) for _ in range(50)] + ["a ab", "ab a", "ab ab", "a ab abc", "abc ab a x", "a b c x",
                                                          "a  b", "h hé", "日 日本 x", "the the, then t"]
    compare_queries(ix, o, queries, [1.0] * n_fields, f"{route} seed={seed}")
    for k, _ in docs[::3]:
        ix.remove_document(k)
        o.remove_document(k)
    compare_queries(ix, o, queries[:30] + queries[-10:], [1.0] * n_fields, f"{route} seed={seed} removed")


@pytest.mark.parametrize("route", ["union_all", "union_cta"])
def test_term_pools_and_consumed_query_terms(monkeypatch, route):
    """zero_to_one.rs:98-121 on nested prefixes: the same expanded term reached from two query terms shares one
    pool of tf uses, a refused entry does not consume its query term, ties keep (query term, expansion) order."""
    _route(monkeypatch, "union_all")
    docs = [
        (0, [["abc"], ["abc abc abd"]]),
        (1, [["abc abd abcd"], ["ab"]]),
        (2, [["ab ab ab"], ["abcd abcd"]]),
        (3, [["abd"], ["abd abd abc abc abc"]]),
        (4, [["x abcde"], ["abcde abc ab a"]]),
        (5, [["a a a a"], ["a"]]),
        (6, [["abc abc"], [""]]),
    ]
    ix, o = both(docs, 2)
    queries = ["ab ab", "ab abc", "abc ab", "a ab", "a a", "ab abd abc", "abc abc", "a ab abc", "abcd ab", "ab x", "abc abd",
               "a abcde", "ab ab x", "abd ab", "a b"]
    compare_queries(ix, o, queries, [1.0, 1.0], "pools")
    fq = FlatQueries.from_strings(queries, TOK)
    b = DeviceBatch(ix, fq, score.zero_to_one.new(), [1.0, 1.0], top_k=5)
    b.run()
    assert b.stats()["union_queries"] >= 10           # the dense route really ran


def test_queries_outside_the_dense_envelope_take_the_list_route(monkeypatch):
    """More than 4 live query terms, or more than 6 candidate slots (4 identical terms need 16), are not class U."""
    _route(monkeypatch, "union_all")
    docs = [(k, [[" ".join(random.Random(k).choice(H.WORDS) for _ in range(5))], ["ab abc a b"]]) for k in range(40)]
    ix, o = both(docs, 2)
    queries = ["a b c x q", "ab ab ab ab", "a ab abc abcd", "a b", "ab abc"]
    compare_queries(ix, o, queries, [1.0, 1.0], "envelope")
    fq = FlatQueries.from_strings(queries, TOK)
    b = DeviceBatch(ix, fq, score.zero_to_one.new(), [1.0, 1.0], top_k=5)
    b.run()
    assert b.stats()["union_queries"] == 2


@pytest.mark.parametrize("route,removed", [("default", False), ("union_all", False), ("union_all", True), ("union_cta", True), ("union_off", False)])
def test_scaled_cfg2_counts_digests_topk(monkeypatch, route, removed):
    _route(monkeypatch, route)
    cfg = W.CONFIGS["cfg2"]
    wl = W.Workload(cfg, n_docs=40_000, vocab=1 << 12)
    ix, o = Index(cfg.n_fields), orc.OracleIndex(cfg.n_fields)
    wl.build_into(ix)
    wl.build_into(o)
    if removed:
        for d in W.Workload(W.CONFIGS["cfg4"], n_docs=40_000, vocab=1 << 12).removed_ordinals():
            ix.remove_document(int(d))
            o.remove_document(int(d))
    fq = wl.queries(160)
    k = 10
    batch = DeviceBatch(ix, fq, score.zero_to_one.new(), cfg.boosts, top_k=k)
    for _ in range(2):                                  # the second run starts from the state the first one left
        batch.run()
        got, st = batch.fetch(), batch.stats()
        exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.ZERO_TO_ONE, cfg.boosts, k,
                                 n_threads=8)
        np.testing.assert_array_equal(got.n_results, exp["n_results"])
        np.testing.assert_array_equal(got.doc_digest, exp["doc_digest"])
        np.testing.assert_array_equal(got.score_digest, exp["score_digest"])
        np.testing.assert_array_equal(got.topk_n, exp["topk_n"])
        for q in range(fq.n_queries):
            n = int(got.topk_n[q])
            np.testing.assert_array_equal(got.topk_doc[q, :n], exp["topk_key"][q, :n].astype(np.uint32))
            np.testing.assert_array_equal(got.topk_score[q, :n], exp["topk_score"][q, :n])
        assert st["pointer_visits"] == exp["score_calls"]
        assert st["results_emitted"] == int(exp["n_results"].sum())
        if route == "union_off":
            assert st["union_queries"] == 0
        else:
            assert st["union_queries"] > 0 and st["rows_streamed_union"] > 0 and st["ms_union"] > 0


def test_full_result_sets_through_the_union_route(monkeypatch):
    _route(monkeypatch, "union_all")
    cfg, ix, o, fq, scorer = _scaled("cfg2", 20_000, 1 << 12, 30)
    qi, docs, scores = ix.query_full_flat(fq, score.zero_to_one.new(), cfg.boosts)
    for q in range(fq.n_queries):
        exp = o.query_tokens(fq.terms_of(q), scorer, cfg.boosts)
        sel = qi == q
        H.assert_same_results([(int(d), float(s)) for d, s in zip(docs[sel], scores[sel])], exp, ctx=f"q={q}")
