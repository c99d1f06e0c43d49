"""-m gpu: incremental image maintenance.  Documents added behind a resident image become a
small DELTA segment — the current trie with only the new documents' posting rows — instead of forcing a full
re-flatten + upload; each segment learns the other's per-term live counts so BM25's idf is the whole index's;
results of the two segments are disjoint by document and merge exactly.  Everything is compared with the CPU
oracle (which, like the reference, has ONE index) bit for bit."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import DeviceBatch, FlatQueries, Index, capi, score
from tests import helpers as H
from tests.test_gpu_parity import TOK, both, compare_queries

pytestmark = pytest.mark.gpu
QUERIES = ["a", "ab", "abc", "ab abc", "a b", "xyz q", "abd  abd", "the then", "zz", "c ca", "b ba bab", "x xy xyz",
           "oy oysters", "hé", "日", "nomatch a", "ab abc abcd a", "q"]


def _add(ix, o, docs, n_fields):
    for key, fields in docs:
        ix.add_document([(lambda d, i=i: d[i]) for i in range(n_fields)], TOK, key, fields)
        o.add_document(key, fields)


@pytest.mark.parametrize("seed", range(4))
def test_add_query_add_query_without_a_rebuild(seed):
    rng = random.Random(9100 + seed)
    n_fields = [1, 2, 2, 3][seed]
    base = H.random_corpus(rng, 60, n_fields, multi_value=(seed == 3))
    ix, o = both(base, n_fields)
    queries = QUERIES + [H.random_query(rng) for _ in range(25)]
    boosts = [1.0] * n_fields
    compare_queries(ix, o, queries, boosts, "base")
    assert ix.n_segments == 1
    main_handle = ix._ix.value
    key = 1000
    for rnd in range(3):                                   # add -> query -> add -> query: the main image is never rebuilt
        more = [(key + k, f) for k, (_, f) in enumerate(H.random_corpus(rng, rng.randint(1, 25), n_fields))]
        key += 100
        _add(ix, o, more, n_fields)
        compare_queries(ix, o, queries, boosts, f"delta round {rnd}")
        assert ix.n_segments == 2 and ix._ix.value == main_handle
        compare_queries(ix, o, queries[:12], [rng.choice([2.0, 0.5, -1.0, 0.0]) for _ in range(n_fields)], f"delta round {rnd} boosts")
    # removals in both segments (lazy, pre-vacuum): live state only, still two segments, still the same main image
    for k in [base[3][0], base[10][0], 1000, 1101]:
        ix.remove_document(k)
        o.remove_document(k)
    compare_queries(ix, o, queries, boosts, "removed in both segments")
    assert ix.n_segments == 2 and ix._ix.value == main_handle
    ix.vacuum()
    o.vacuum()
    compare_queries(ix, o, queries, boosts, "vacuumed")    # vacuum prunes the trie: one fresh image
    assert ix.n_segments == 1
    _add(ix, o, [(5000, [["abc zz new"]] * n_fields)], n_fields)
    compare_queries(ix, o, queries + ["new", "ne"], boosts, "after vacuum + add")
    assert ix.n_segments == 2


def test_document_frequency_is_over_both_segments():
    """BM25's idf (bm25.rs:41-56) counts the term's postings in the WHOLE index: a term that is rare in the main
    segment and frequent in the delta must score with the combined frequency in both."""
    docs = [(k, [["common filler"], ["rare" if k == 0 else "filler"]]) for k in range(30)]
    ix, o = both(docs, 2)
    compare_queries(ix, o, ["rare", "common", "ra"], [1.0, 1.0], "before")
    _add(ix, o, [(100 + k, [["rare rare"], ["rare common"]]) for k in range(12)], 2)
    compare_queries(ix, o, ["rare", "common", "ra", "rare common", "filler rare"], [1.0, 1.0], "after")
    assert ix.n_segments == 2
    r = ix.query("rare", score.bm25.new(), TOK, [1.0, 1.0])
    assert {x.key for x in r} == {0} | {100 + k for k in range(12)}


def test_a_grown_delta_is_folded_into_the_main_image(monkeypatch):
    monkeypatch.setattr(Index, "DELTA_MIN_ROWS", 0)
    monkeypatch.setattr(Index, "DELTA_MAX_FRACTION", 0.5)
    rng = random.Random(77)
    base = H.random_corpus(rng, 40, 2)
    ix, o = both(base, 2)
    compare_queries(ix, o, QUERIES[:8], [1.0, 1.0], "base")
    _add(ix, o, [(500 + k, f) for k, (_, f) in enumerate(H.random_corpus(rng, 5, 2))], 2)
    compare_queries(ix, o, QUERIES[:8], [1.0, 1.0], "small delta")
    assert ix.n_segments == 2
    _add(ix, o, [(600 + k, f) for k, (_, f) in enumerate(H.random_corpus(rng, 60, 2))], 2)
    compare_queries(ix, o, QUERIES, [1.0, 1.0], "grown delta")
    assert ix.n_segments == 1                               # more than half of the main image's rows: compacted


def test_staged_batches_and_images_see_one_segment():
    rng = random.Random(5)
    base = H.random_corpus(rng, 50, 2)
    ix, o = both(base, 2)
    fq = FlatQueries.from_strings(QUERIES, TOK)
    b = DeviceBatch(ix, fq, score.bm25.new(), [1.0, 1.0], top_k=5)
    b.run()
    ix.remove_document(base[0][0]); o.remove_document(base[0][0])
    b.run()                                                 # removals follow a staged batch (live state only)
    got = b.fetch()
    for q, text in enumerate(QUERIES):
        assert int(got.n_results[q]) == len(o.query(text, orc.BM25, [1.0, 1.0]))
    _add(ix, o, [(900, [["abc new"], ["abd"]])], 2)
    with pytest.raises(capi.ProblyError):
        b.run()                                             # documents were added: this batch is stale, loudly
    b2 = DeviceBatch(ix, fq, score.bm25.new(), [1.0, 1.0], top_k=5)      # staging compacts
    assert ix.n_segments == 1
    b2.run()
    got = b2.fetch()
    for q, text in enumerate(QUERIES):
        assert int(got.n_results[q]) == len(o.query(text, orc.BM25, [1.0, 1.0]))
    assert ix.expand_term("ab") == o.expand_term("ab")


@pytest.mark.parametrize("layout", ["auto", "wide"])
def test_device_side_flatten_equals_the_host_flattened_image(monkeypatch, layout):
    """pb_index_create_from_builder sorts the builder's (term, doc, tf) log on the device and writes the
    tile-blocked posting columns itself.  The image it builds must answer exactly like the one uploaded from the
    host-flattened pb_index_image (same layout decision, same rows in the same order)."""
    import ctypes as C
    from probly_search_b200 import workload as W
    monkeypatch.setenv("PB_POSTING_LAYOUT", layout)
    cfg = W.CONFIGS["cfg2"]
    wl = W.Workload(cfg, n_docs=30_000, vocab=1 << 12)
    ix = Index(cfg.n_fields)
    wl.build_into(ix)
    for d in range(0, 30_000, 17):
        ix.remove_document(d)
    L = capi.lib()
    ix.sync_device()                                   # device-side flatten (the default path of the host mirror)
    im = ix.flatten()                                  # the host-flattened image with its posting columns
    h = C.c_void_p()
    capi.check(L.pb_index_create(C.byref(im), 0, C.byref(h)))
    la, lb = capi.DeviceLayout(), capi.DeviceLayout()
    capi.check(L.pb_index_device_layout(ix._ix, C.byref(la)))
    capi.check(L.pb_index_device_layout(h, C.byref(lb)))
    assert (la.narrow, la.bytes_per_row, la.posting_bytes, list(la.fl_bits)) == (lb.narrow, lb.bytes_per_row, lb.posting_bytes, list(lb.fl_bits))
    assert bool(la.narrow) == (layout == "auto")
    from probly_search_b200.index import BatchResults
    for scorer, mode in ((score.bm25.new(), 0), (score.zero_to_one.new(), 1)):
        fq = wl.queries(300, mode=mode)
        d, _keep = ix._desc(fq, scorer, cfg.boosts, 10)
        outs = []
        for handle in (ix._ix, h):
            r = BatchResults(fq.n_queries, 10)
            rs = r.c_struct()
            capi.check(L.pb_query_batch(handle, C.byref(d), C.byref(rs)))
            outs.append(r)
        for name in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score"):
            np.testing.assert_array_equal(getattr(outs[0], name), getattr(outs[1], name), err_msg=name)
        assert int(outs[0].n_results.sum()) > 0
    L.pb_index_destroy(h)
