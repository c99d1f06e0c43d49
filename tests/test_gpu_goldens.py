"""-m gpu: the reference's own golden vectors through the CUDA path, called via
the C ABI behind the `Index::query` mirror.  Reads like the reference's tests on purpose."""
import pytest

from probly_search_b200 import Index, score
from tests import helpers as H
from tests.golden import reference_cases as G

pytestmark = pytest.mark.gpu


def tokenizer(s):                      # src/lib.rs:42-44
    return s.split(" ")


def calc(name):
    return score.bm25.new() if name == G.BM25 else score.zero_to_one.new()


def build(case):
    n = case["fields"]
    ix = Index(n)
    for key, texts in case["docs"]:
        ix.add_document([(lambda d, i=i: [d[i]]) for i in range(n)], tokenizer, key, texts)
    return ix


def run_ops(ix, ops):
    for op in ops:
        if op[0] == "remove":
            ix.remove_document(op[1])
        elif op[0] == "vacuum":
            ix.vacuum()
        else:
            _, q, scorer, boosts, expected, exact = op
            got = [(r.key, r.score) for r in ix.query(q, calc(scorer), tokenizer, boosts)]
            H.assert_same_results(got, expected, ctx=q, tol=1e-8 if not exact else 1e-9, exact=exact)
            # the batch entry point must agree with the full result set
            (top,) = ix.query_batch([q], calc(scorer), tokenizer, boosts, top_k=3)
            exp_top = sorted(expected, key=lambda r: (-r[1], r[0]))[:3]
            assert [(r.key) for r in top] == [k for k, _ in exp_top] or len({s for _, s in exp_top}) < len(exp_top)
            assert [r.score for r in top] == pytest.approx([s for _, s in exp_top], abs=1e-8)


@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_reference_goldens(case):
    run_ops(build(case), case["ops"])


@pytest.mark.parametrize("case", G.DERIVED_CASES, ids=[c["name"] for c in G.DERIVED_CASES])
def test_derived_goldens(case):
    run_ops(build(case), case["ops"])


@pytest.mark.parametrize("case", G.EXPANSION_CASES, ids=[c["name"] for c in G.EXPANSION_CASES])
def test_expansion_order(case):
    assert build(case).expand_term(case["term"]) == case["expected"]


def test_df_clamp_does_not_fail():      # tests/document_frequency.rs:5-32
    c = G.DF_CLAMP_CASE
    ix = build(c)
    got = ix.query(c["query"], score.bm25.new(), tokenizer, [1.0])
    assert len(got) == 1 and got[0].key == 0 and got[0].score > 0


def test_custom_calculator_is_rejected():
    class Mine(score.ScoreCalculator):
        pass
    ix = build(G.CASES[0])
    with pytest.raises(TypeError):
        ix.query("a", Mine(), tokenizer, [1.0])
