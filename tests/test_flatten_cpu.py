"""CPU-only: the product's host index builder + flattener (csrc/builder.cpp) against the oracle's
faithful linked structures — trie shape, expansion order, de-duplicated postings, multiplicities,
field statistics, removal and vacuum.  No compute entry point is called (no GPU here)."""
import math
import random

import numpy as np
import pytest

from oracle import oracle as orc
from probly_search_b200 import Index
from tests import helpers as H
from tests.golden import reference_cases as G

TOK = orc.tokenizer


def build_pair(docs, n_fields):
    ix, o = Index(n_fields), orc.OracleIndex(n_fields)
    for key, fields in docs:
        accessors = [(lambda d, i=i: d[i]) for i in range(n_fields)]
        ix.add_document(accessors, TOK, key, fields)
        o.add_document(key, fields)
    return ix, o


def check_against_oracle(ix: Index, o, n_fields, prefixes):
    info = ix.info()
    assert info.n_live_docs == o.docs_len()
    assert info.n_nodes == o.count_nodes()
    assert info.n_pointers == o.arena_doc_len() or info.n_removed_pending > 0
    stats = o.field_stats()
    for f in range(n_fields):
        assert info.field_sum[f] == stats[f][0]
        a, b = info.field_avg[f], stats[f][1]
        assert (math.isnan(a) and math.isnan(b)) or a == b
    im = ix.flatten()
    a = H.image_arrays(im)
    assert im.n_rows_padded % 128 == 0 and im.n_rows_padded >= im.n_rows + 128
    # expansion order = the reference's DFS / prepend order (query.rs:109-147)
    for p in prefixes:
        assert H.image_expand(a, p) == o.expand_term(p), p
    # postings: one row per (term, doc), docs ascending, tf vector and multiplicity preserved
    id2key = ix._id_to_key
    for t in range(int(im.n_terms)):
        term = H.image_term_string(a, t)
        assert int(a["term_byte_len"][t]) == len(term.encode("utf-8"))
        r0, r1 = int(a["term_row_begin"][t]), int(a["term_row_begin"][t + 1])
        docs = a["post_doc"][r0:r1]
        assert np.all(np.diff(docs.astype(np.int64)) > 0)
        ptrs = o.postings(term)
        exp = {}
        for key, tf in ptrs:
            exp.setdefault(key, []).append(tf)
        got_keys = [id2key[int(a["doc_key"][d])] for d in docs]
        assert sorted(got_keys) == sorted(exp.keys()), term
        for i, key in enumerate(got_keys):
            tf = [int(a["post_tf"][f][r0 + i]) for f in range(n_fields)]
            assert all(tf == e for e in exp[key]), (term, key)
            assert len(exp[key]) == sum(tf)          # multiplicity = sum of tf (SURVEY §3.4 rule 1)
    return a


@pytest.mark.parametrize("seed", range(6))
def test_random_corpora_match_oracle_structure(seed):
    rng = random.Random(seed)
    n_fields = rng.choice([1, 2, 3])
    docs = H.random_corpus(rng, rng.randint(1, 40), n_fields, multi_value=(seed % 2 == 0))
    ix, o = build_pair(docs, n_fields)
    prefixes = ["", "a", "ab", "abc", "b", "x", "z", "h", "hé", "日", "nomatch", "t", "the"]
    check_against_oracle(ix, o, n_fields, prefixes)
    # remove a third of the docs: lazily (mask only), then vacuum
    victims = [k for k, _ in docs if rng.random() < 0.33]
    for k in victims:
        ix.remove_document(k)
        o.remove_document(k)
    info = ix.info()
    assert info.n_removed_pending == len(victims)
    im = ix.flatten()
    a = H.image_arrays(im)
    removed = {ix._id_to_key[int(a["doc_key"][d])] for d in range(int(im.n_docs))
               if (int(a["removed"][d >> 5]) >> (d & 31)) & 1}
    assert removed == set(victims)
    stats = o.field_stats()
    for f in range(n_fields):
        x, y = info.field_avg[f], stats[f][1]
        assert info.field_sum[f] == stats[f][0] and ((math.isnan(x) and math.isnan(y)) or x == y)
    ix.vacuum()
    o.vacuum()
    check_against_oracle(ix, o, n_fields, prefixes)
    # add more documents after the vacuum: re-created nodes must come FIRST among their siblings
    more = H.random_corpus(rng, 10, n_fields)
    for key, fields in more:
        accessors = [(lambda d, i=i: d[i]) for i in range(n_fields)]
        ix.add_document(accessors, TOK, 1000 + key, fields)
        o.add_document(1000 + key, fields)
    check_against_oracle(ix, o, n_fields, prefixes)


@pytest.mark.parametrize("case", G.EXPANSION_CASES, ids=[c["name"] for c in G.EXPANSION_CASES])
def test_expansion_goldens_on_image(case):
    ix = Index(case["fields"])
    for key, texts in case["docs"]:
        ix.add_document([(lambda d, i=i: [d[i]]) for i in range(case["fields"])], TOK, key, texts)
    a = H.image_arrays(ix.flatten())
    assert H.image_expand(a, case["term"]) == case["expected"]


def test_reference_structure_pins():
    # src/index.rs:496-545, 606-617, 620-658, 738-783 restated on the builder
    ix = Index(1)
    ix.add_document([lambda d: [d]], TOK, 1, "a b c")
    i = ix.info()
    assert (i.n_live_docs, i.field_sum[0], i.field_avg[0], i.n_nodes) == (1, 3, 3.0, 4)
    a = H.image_arrays(ix.flatten())
    assert H.image_expand(a, "") == ["c", "b", "a"]          # most recently created child first
    ix2 = Index(1)
    ix2.add_document([lambda d: [d]], TOK, 1, "a  b")          # empty token ignored
    assert ix2.info().field_sum[0] == 2
    ix3 = Index(1)
    ix3.add_document([lambda d: [d]], TOK, 1, "a")
    ix3.remove_document(1)
    ix3.vacuum()
    i3 = ix3.info()
    assert i3.n_live_docs == 0 and i3.field_sum[0] == 0 and math.isnan(i3.field_avg[0])
    assert i3.n_nodes == 1 and i3.n_rows == 0
    ix4 = Index(1)
    ix4.add_document([lambda d: [d]], TOK, 1, "ab cd")
    ix4.add_document([lambda d: [d]], TOK, 2, "ab ef")
    assert ix4.info().n_nodes == 7


def test_builder_rejects_bad_input():
    from probly_search_b200 import capi
    ix = Index(2)
    ix.add_document([lambda d: ["x"], lambda d: ["y"]], TOK, 1, None)
    with pytest.raises(capi.ProblyError) as e:
        ix.add_document([lambda d: ["x"], lambda d: ["y"]], TOK, 1, None)
    assert e.value.code == capi.PB_ERR_DUPLICATE_KEY
    with pytest.raises(capi.ProblyError):
        Index(9)


def test_synthetic_corpus_matches_oracle_counts():
    from probly_search_b200 import workload as W
    wl = W.Workload(W.CONFIGS["cfg1"], n_docs=3000, vocab=1 << 10)
    ix, o = Index(2), orc.OracleIndex(2)
    wl.build_into(ix)
    wl.build_into(o)
    i = ix.info()
    assert i.n_live_docs == o.docs_len() == 3000
    assert i.n_nodes == o.count_nodes()
    assert i.n_pointers == o.arena_doc_len()
    assert [(i.field_sum[f], i.field_avg[f]) for f in range(2)] == o.field_stats()
    a = H.image_arrays(ix.flatten())
    for p in ["a", "ab", "s", "sz", "q"]:
        assert H.image_expand(a, p) == o.expand_term(p)


def test_remove_after_flatten_updates_only_the_live_state():
    """remove_document is lazy (src/index.rs:161-191): the flattened structure stays, the image's removed bitmap /
    n_live_docs / field averages follow in O(1) — no second O(rows) flatten (round-1 advisory)."""
    import ctypes as C
    from probly_search_b200 import Index
    tok = lambda s: s.split(" ")
    ix = Index(1)
    for k, d in enumerate(["a b c", "a b", "c c d", "e"]):
        ix.add_document([lambda d: [d]], tok, k, d)
    im0 = ix.flatten()
    p_blocks = C.addressof(im0.post_blocks.contents)
    rows0, avg0 = int(im0.n_rows), float(im0.field_avg[0])
    ix.remove_document(2)
    im1 = ix.flatten()
    assert C.addressof(im1.post_blocks.contents) == p_blocks and int(im1.n_rows) == rows0      # same buffers: not re-flattened
    assert int(im1.n_removed) == 1 and int(im1.n_live_docs) == 3
    assert (im1.removed_bitmap[0] >> 2) & 1
    assert float(im1.field_avg[0]) == (3 + 2 + 1) / 3 and avg0 == (3 + 2 + 3 + 1) / 4
    ix.vacuum()
    im2 = ix.flatten()
    assert int(im2.n_removed) == 0 and im2.removed_bitmap[0] == 0 and int(im2.n_docs) == 4    # the ordinal stays, unmasked: it owns no row
    assert int(im2.n_rows) == rows0 - 2
