"""Mirror of `probly_search::score` (src/score/mod.rs, src/score/calculator.rs:33-70).

The reference's `ScoreCalculator` trait is arbitrary user code called per posting
(before_each / score / finalize).  On the GPU only the two built-in calculators exist as
device code, so the trait is kept as a *marker*: `bm25.new()` and `zero_to_one.new()` return
objects the query path recognises; any other implementor is rejected — there is no CPU
fallback (north_star).
"""
from __future__ import annotations

from . import capi


class ScoreCalculator:
    """Marker base (calculator.rs:33).  `_pb_scorer` selects the device implementation."""
    _pb_scorer: int | None = None


class BM25(ScoreCalculator):
    """src/score/default/bm25.rs:14-26 — public fields `bm25k1`, `bm25b`."""
    _pb_scorer = capi.PB_SCORER_BM25

    def __init__(self, bm25k1: float = 1.2, bm25b: float = 0.75):
        self.bm25k1 = bm25k1
        self.bm25b = bm25b


class ZeroToOne(ScoreCalculator):
    """src/score/default/zero_to_one.rs:24-39.  Stateless here: the per-query state the
    reference keeps in `score_by_document_and_field` lives in the device workspace."""
    _pb_scorer = capi.PB_SCORER_ZERO_TO_ONE


class _Bm25Module:
    BM25 = BM25

    @staticmethod
    def new() -> BM25:          # bm25.rs:21-26
        return BM25()


class _ZeroToOneModule:
    ZeroToOne = ZeroToOne

    @staticmethod
    def new() -> ZeroToOne:     # zero_to_one.rs:35-39
        return ZeroToOne()


bm25 = _Bm25Module()
zero_to_one = _ZeroToOneModule()


def scorer_params(calc) -> tuple[int, float, float]:
    if not isinstance(calc, ScoreCalculator) or calc._pb_scorer is None:
        raise TypeError(
            f"{type(calc).__name__} is not a device score calculator: only score.bm25.new() and "
            "score.zero_to_one.new() can run on the GPU and this library has no CPU fallback")
    if isinstance(calc, BM25):
        return calc._pb_scorer, float(calc.bm25k1), float(calc.bm25b)
    return calc._pb_scorer, 1.2, 0.75
