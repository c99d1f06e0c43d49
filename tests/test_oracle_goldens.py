"""Pins the CPU oracle (oracle/probly_oracle.cpp) against every golden vector the
reference's own tests hold for the path — CPU only."""
import math

import pytest

from oracle import oracle as orc
from tests.golden import reference_cases as G

SCORER = {G.BM25: orc.BM25, G.Z2O: orc.ZERO_TO_ONE}


def build(case):
    ix = orc.OracleIndex(case["fields"])
    for key, texts in case["docs"]:
        ix.add_document(key, [[t] for t in texts])
    return ix


def run_ops(ix, ops):
    for op in ops:
        if op[0] == "remove":
            ix.remove_document(op[1])
        elif op[0] == "vacuum":
            ix.vacuum()
        else:
            _, q, scorer, boosts, expected, exact = op
            got = ix.query(q, SCORER[scorer], boosts)
            assert len(got) == len(expected), (q, got, expected)
            for (gk, gs), (ek, es) in zip(got, expected):
                assert gk == ek, (q, got, expected)
                if exact:
                    assert gs == es, (q, got, expected)      # reference asserts f64 ==
                assert abs(gs - es) < 1e-8


@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_reference_goldens(case):
    run_ops(build(case), case["ops"])


@pytest.mark.parametrize("case", G.DERIVED_CASES, ids=[c["name"] for c in G.DERIVED_CASES])
def test_derived_goldens(case):
    run_ops(build(case), case["ops"])


@pytest.mark.parametrize("case", G.EXPANSION_CASES, ids=[c["name"] for c in G.EXPANSION_CASES])
def test_expansion_order(case):
    assert build(case).expand_term(case["term"]) == case["expected"]


def test_corpus_d_structure():
    ix = build(dict(fields=2, docs=G.CORPUS_D))
    assert ix.field_stats() == G.CORPUS_D_FIELDS
    assert ix.expand_term(G.CORPUS_D_EXPAND["term"]) == G.CORPUS_D_EXPAND["expected"]
    ix.remove_document(1)
    ix.remove_document(3)
    assert ix.field_stats() == G.CORPUS_D_FIELDS_AFTER_REMOVE
    assert ix.docs_len() == 3


def test_df_clamp_does_not_panic():
    c = G.DF_CLAMP_CASE
    ix = build(c)
    got = ix.query(c["query"], orc.BM25, [1.0])
    # "the" matches (plus expansion "the,"): one doc, finite positive score
    assert len(got) == 1 and got[0][0] == 0 and math.isfinite(got[0][1]) and got[0][1] > 0


# ---- src/index.rs:492-618 structure pins --------------------------------------------------------
def test_add_one_document_three_terms():            # index.rs:496-545
    ix = orc.OracleIndex(1)
    ix.add_document(1, [["a b c"]])
    assert ix.docs_len() == 1
    assert ix.field_stats() == [(3, 3.0)]
    assert ix.children_chars("") == ["c", "b", "a"]          # children are prepended
    assert ix.children_chars("c") == []
    assert ix.postings("c") == [(1, [1])]


def test_add_shared_terms():                        # index.rs:547-604
    ix = orc.OracleIndex(1)
    ix.add_document(1, [["a b c"]])
    ix.add_document(2, [["b c d"]])
    assert ix.docs_len() == 2
    assert ix.field_stats() == [(6, 3.0)]
    assert ix.children_chars("") == ["d", "c", "b", "a"]
    assert ix.postings("c") == [(2, [1]), (1, [1])]          # postings are prepended


def test_ignore_empty_tokens():                     # index.rs:606-617
    ix = orc.OracleIndex(1)
    ix.add_document(1, [["a  b"]])
    assert ix.field_stats() == [(2, 2.0)]
    assert sorted(ix.children_chars("")) == ["a", "b"]


def test_delete_and_vacuum():                       # index.rs:620-658
    ix = orc.OracleIndex(1)
    assert ix.arena_doc_len() == 0
    ix.add_document(1, [["a"]])
    ix.remove_document(1)
    ix.vacuum()
    assert ix.docs_len() == 0
    (s, a), = ix.field_stats()
    assert s == 0 and math.isnan(a)
    assert ix.children_chars("") == []
    assert ix.arena_doc_len() == 0
    assert ix.arena_index_len() == 1                          # only the root is left


def test_count_nodes():                             # index.rs:738-783
    ix = orc.OracleIndex(1)
    assert ix.count_nodes() == 1
    ix.add_document(1, [["abc"]])
    ix.add_document(1, [["abe"]])
    assert ix.count_nodes() == 5
    ix = orc.OracleIndex(1)
    ix.add_document(1, [["ab cd"]])
    ix.add_document(1, [["ab ef"]])
    assert ix.count_nodes() == 7


def test_duplicate_postings_and_df():               # SURVEY §3.4 rules 1-2
    ix = orc.OracleIndex(2)
    ix.add_document(7, [["x x y"], ["x"]])
    # one pointer per OCCURRENCE, each carrying the full tf vector
    assert ix.postings("x") == [(7, [2, 1])] * 3
    assert ix.postings("y") == [(7, [1, 0])]


def test_multi_valued_field_length_is_last_value():   # SURVEY §3.4 rule 8 (index.rs:112-114)
    ix = orc.OracleIndex(1)
    ix.add_document(0, [["a b c", "d"]])
    assert ix.field_stats() == [(4, 4.0)]
    # field_length = 1 (last value); BM25 on "a" uses fl=1, avg=4
    got = ix.query("a", orc.BM25, [1.0])
    idf = math.log(1.0 + (1 - 1 + 0.5) / (1 + 0.5))
    tf = ((1.2 + 1.0) * 1.0) / (1.2 * ((1.0 - 0.75) + 0.75 * (1.0 / 4.0)) + 1.0)
    assert got == [(0, tf * idf * 1.0 * 1.0)]
