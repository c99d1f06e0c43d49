"""probly_search_b200 — B200-native query hot path of probly-search behind the reference's
`Index::query` / `ScoreCalculator` surface.  All compute happens in hand-written sm_100a CUDA
kernels reached through the C ABI in include/probly_b200.h; there is no CPU fallback."""
from . import score
from .index import BatchResults, DeviceBatch, FlatQueries, Index, QueryResult

__all__ = ["Index", "QueryResult", "FlatQueries", "BatchResults", "DeviceBatch", "score"]
