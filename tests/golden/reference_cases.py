"""Golden vectors transcribed from the reference's own tests.

The reference is a Rust crate that cannot be built here (no rustc/cargo), so these are the
known-answer vectors its `cargo test` asserts, copied as DATA (corpus, query, expected
`(key, score)` list) with the file:line that pins each.  They are used twice: to pin the CPU
oracle (tests/test_oracle_goldens.py, CPU-only) and as known-answer tests of the CUDA path
through the C ABI (tests/test_gpu_goldens.py, -m gpu).

Case format:
  fields : number of indexed fields (Index::new(n))
  docs   : [(key, [field0 text, field1 text, ...])]       added in this order
  ops    : list of steps executed in order:
             ("query", query, scorer, boosts, expected, exact)  expected in (score desc, key asc) order
             ("remove", key) / ("vacuum",)
  `exact` True  -> the reference asserts f64 `==` (lib.rs:63, integrations_tests.rs)
          False -> the reference asserts |a-b| < 1e-8 (query.rs:172-176)
"""

BM25 = "bm25"
Z2O = "zero_to_one"

_TWO = [(1, ["a b c", "hello world"]), (2, ["c d e", "lorem ipsum"])]


def _titles(titles):
    # build_test_index (lib.rs:72-83): 1 field, ids 0..n
    return [(i, [t]) for i, t in enumerate(titles)]


CASES = [
    # ---- src/score/default/bm25.rs:104-136 (exact ==, 1 field) -------------------------------
    dict(name="bm25_unit_doc1", ref="src/score/default/bm25.rs:104-116", fields=1,
         docs=_titles(["a b c", "c d e"]),
         ops=[("query", "a", BM25, [1.0], [(0, 0.6931471805599453)], True)]),
    dict(name="bm25_unit_doc1_and_2", ref="src/score/default/bm25.rs:118-136", fields=1,
         docs=_titles(["a b c", "c d e"]),
         ops=[("query", "c", BM25, [1.0], [(0, 0.1823215567939546), (1, 0.1823215567939546)], True)]),
    # ---- src/query.rs:181-339 (1e-8, 2 fields) ------------------------------------------------
    dict(name="query_doc1", ref="src/query.rs:181-211", fields=2, docs=_TWO,
         ops=[("query", "a", BM25, [1.0, 1.0], [(1, 0.6931471805599453)], False)]),
    dict(name="query_doc1_and_2", ref="src/query.rs:213-258", fields=2, docs=_TWO,
         ops=[("query", "c", BM25, [1.0, 1.0], [(1, 0.1823215567939546), (2, 0.1823215567939546)], False)]),
    dict(name="query_expanding", ref="src/query.rs:260-292", fields=2, docs=_TWO,
         ops=[("query", "h", BM25, [1.0, 1.0], [(1, 0.12637567304702957)], False)]),
    dict(name="query_disjunction", ref="src/query.rs:294-338", fields=2, docs=_TWO,
         ops=[("query", "a d", BM25, [1.0, 1.0], [(1, 0.6931471805599453), (2, 0.6931471805599453)], False)]),
    # ---- tests/integrations_tests.rs:27-93 (exact ==) -----------------------------------------
    dict(name="integration_bm25", ref="tests/integrations_tests.rs:27-93", fields=2,
         docs=[(0, ["abc", "dfg"]), (1, ["dfgh", "abcd"])],
         ops=[("query", "abc", BM25, [1.0, 1.0], [(0, 0.6931471805599453), (1, 0.28104699650060755)], True),
              ("remove", 0), ("vacuum",),
              ("query", "abc", BM25, [1.0, 1.0], [(1, 0.1166450426074421)], True)]),
    # ---- tests/integrations_tests.rs:95-149 (exact ==, remove WITHOUT vacuum) ------------------
    dict(name="integration_zero_to_one", ref="tests/integrations_tests.rs:95-149", fields=2,
         docs=[(0, ["abc", "dfg"]), (1, ["dfgh", "abcd"])],
         ops=[("query", "abc", Z2O, [1.0, 1.0], [(0, 1.0), (1, 0.75)], True),
              ("remove", 0),
              ("query", "abc", Z2O, [1.0, 1.0], [(1, 0.75)], True)]),
    # ---- src/score/default/zero_to_one.rs:138-404 (exact ==) -----------------------------------
    dict(name="z2o_partial_matching", ref="src/score/default/zero_to_one.rs:138-156", fields=1,
         docs=_titles(["abc", "abcefg", "abcefghij"]),
         ops=[("query", "abc", Z2O, [1.0], [(0, 1.0), (1, 0.5), (2, 0.33333333333333337)], True)]),
    dict(name="z2o_partial_matching_repeating", ref="src/score/default/zero_to_one.rs:158-170", fields=1,
         docs=_titles(["abcdef abcdefghi"]),
         ops=[("query", "abc abc", Z2O, [1.0], [(0, 0.4166666666666667)], True)]),
    dict(name="z2o_penalize_repeating_query_terms", ref="src/score/default/zero_to_one.rs:172-181", fields=1,
         docs=_titles(["abc"]),
         ops=[("query", "abc abc", Z2O, [1.0], [(0, 0.5)], True)]),
    dict(name="z2o_penalize_missing_repeating", ref="src/score/default/zero_to_one.rs:183-192", fields=1,
         docs=_titles(["abc abc"]),
         ops=[("query", "abc", Z2O, [1.0], [(0, 0.5)], True)]),
    dict(name="z2o_bounded_by_one", ref="src/score/default/zero_to_one.rs:193-205", fields=1,
         docs=_titles(["abc abc"]),
         ops=[("query", "abc ab", Z2O, [1.0], [(0, 0.8333333333333334)], True)]),
    dict(name="z2o_bounded_by_one_2", ref="src/score/default/zero_to_one.rs:207-216", fields=1,
         docs=_titles(["abc ab"]),
         ops=[("query", "abc abc", Z2O, [1.0], [(0, 0.5)], True)]),
    dict(name="z2o_bounded_be_one", ref="src/score/default/zero_to_one.rs:218-230", fields=1,
         docs=_titles(["oy oy oysters"]),
         ops=[("query", "oy oy oysters", Z2O, [1.0], [(0, 1.0)], True)]),
    dict(name="z2o_multiple_results", ref="src/score/default/zero_to_one.rs:232-265", fields=1,
         docs=_titles(["abcdef", "abc abcdef", "abcdef abcdef", "abcdef abcdefghi", "def abcdef"]),
         ops=[("query", "abc", Z2O, [1.0], [(0, 0.5), (1, 0.5), (2, 0.25), (3, 0.25), (4, 0.25)], True)]),
    dict(name="z2o_multiple_results_penalize", ref="src/score/default/zero_to_one.rs:267-307", fields=1,
         docs=_titles(["abcdef", "abc abcdef", "abcdef abcdef", "abcdef abcdefghi", "def abcdef"]),
         ops=[("query", "abc abc", Z2O, [1.0],
               [(1, 0.75), (2, 0.5), (3, 0.4166666666666667), (0, 0.25), (4, 0.25)], True)]),
    dict(name="z2o_multi_field", ref="src/score/default/zero_to_one.rs:309-356", fields=2,
         docs=[(0, ["abc", "abc"]), (1, ["abcefg", "abcefg"]), (2, ["abcefghij", "abcefghij"])],
         ops=[("query", "abc", Z2O, [1.0, 1.0], [(0, 1.0), (1, 0.5), (2, 0.33333333333333337)], True)]),
    dict(name="z2o_multi_field_ignore_lowest", ref="src/score/default/zero_to_one.rs:358-404", fields=2,
         docs=[(0, ["abc", "a"]), (1, ["abcefg", "a"]), (2, ["abcefghij", "a"])],
         ops=[("query", "abc", Z2O, [1.0, 1.0], [(0, 1.0), (1, 0.5), (2, 0.33333333333333337)], True)]),
]

# ---- src/query.rs:343-387 — expansion order (DFS, most-recently-created child first) -----------
EXPANSION_CASES = [
    dict(name="expand_all", ref="src/query.rs:343-364", fields=2,
         docs=[(1, ["abc", "hello world"]), (2, ["adef", "lorem ipsum"])],
         term="a", expected=["adef", "abc"]),
    dict(name="expand_none", ref="src/query.rs:366-387", fields=2,
         docs=[(1, ["abc def", "hello world"]), (2, ["adef abc", "lorem ipsum"])],
         term="x", expected=[]),
]

# ---- tests/document_frequency.rs:5-32 — must not panic when pointer-count df > docs.len() ------
DF_CLAMP_CASE = dict(
    name="df_clamp", ref="tests/document_frequency.rs:5-32", fields=1,
    docs=[(0, ["this is text with lots of the, the, the, the"])],
    query="What did the author do growing up?")

# ---- SURVEY.md Appendix C — DERIVED (not reference-pinned) known answers ------------------------
# Produced by the survey's faithful restatement (which reproduced every golden above bit-exactly)
# for the cases no reference test pins: boosts != 1, removed-but-not-vacuumed docs, the
# max-merger's multi-term/multi-expansion interaction, repeated terms + empty tokens.
CORPUS_D = [
    (0, ["abc abd", "abc xyz abd"]),
    (1, ["ab", "abcde abc"]),
    (2, ["xyz", "ab abd abd"]),
    (3, ["abd abc", "q"]),
    (4, ["zz", "abcde"]),
]
DERIVED_CASES = [
    dict(name="D_boosts_prefix_multi_removed", ref="SURVEY.md Appendix C (derived)", fields=2, docs=CORPUS_D,
         ops=[
             ("query", "abc", BM25, [2.0, 0.5],
              [(0, 0.6089515441682334), (3, 0.48953634428258846), (4, 0.1583099010294772), (1, 0.14384103622589042)], True),
             ("query", "abc", Z2O, [2.0, 0.5], [(4, 0.6), (0, 0.5), (1, 0.5), (3, 0.5)], True),
             ("query", "ab abc", BM25, [1.0, 1.0],
              [(1, 1.2790216721025205), (2, 0.7268042347843698), (0, 0.6796809191540738),
               (4, 0.5622092002640836), (3, 0.3440131255200019)], True),
             ("query", "ab abc", Z2O, [1.0, 1.0],
              [(0, 0.8333333333333334), (3, 0.8333333333333334), (1, 0.7), (2, 0.3333333333333333), (4, 0.3)], True),
             ("query", "abd  abd", BM25, [1.0, 1.0],
              [(0, 0.29253527891895525), (2, 0.20978085411198397), (3, 0.14806355863428702)], True),
             ("query", "abd  abd", Z2O, [1.0, 1.0],
              [(2, 0.6666666666666666), (0, 0.3333333333333333), (3, 0.3333333333333333)], True),
             ("remove", 1), ("remove", 3),
             ("query", "ab", BM25, [2.0, 0.5],
              [(2, 0.4390921655924589), (0, 0.40173157966456696), (4, 0.14281915806561288)], True),
             ("query", "ab", Z2O, [2.0, 0.5],
              [(4, 0.4), (0, 0.33333333333333337), (2, 0.3333333333333333)], True),
             ("query", "abc xyz", BM25, [1.0, 1.0],
              [(0, 1.2318260980626499), (2, 0.5235483465015789), (4, 0.3682518373141765)], True),
             ("query", "abc xyz", Z2O, [1.0, 1.0], [(0, 0.6666666666666666), (2, 0.5), (4, 0.3)], True),
         ]),
]
CORPUS_D_EXPAND = dict(term="ab", expected=["ab", "abd", "abc", "abcde"])
CORPUS_D_FIELDS = [(7, 1.4), (10, 2.0)]
CORPUS_D_FIELDS_AFTER_REMOVE = [(4, 1.3333333333333333), (7, 2.3333333333333335)]
