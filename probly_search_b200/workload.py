"""Synthetic corpora and query sets of SURVEY.md §8(d) (generator: csrc/workload.cpp).

CONFIGS mirrors BASELINE.json `configs` (index = cfg number).  Sizes can be scaled down with
`scale` for tests; the generator is deterministic in (seed, doc index), so a prefix of a
corpus is the same corpus at a smaller size.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Iterator, List, Sequence, Tuple

import numpy as np

from .index import FlatQueries

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_lib", "libprobly_workload.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            from . import build as _build
            _build.build()
        L = C.CDLL(_SO)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.wl_new.restype = vp
        L.wl_new.argtypes = [u64, u32, u32, vp, vp]
        L.wl_free.argtypes = [vp]
        L.wl_vocab_word.restype = u32
        L.wl_vocab_word.argtypes = [vp, u32, vp, u32]
        L.wl_doc_bounds.argtypes = [vp, u64, vp, vp]
        L.wl_gen_docs.argtypes = [vp, u64, u64, vp, vp, vp, vp, vp]
        L.wl_gen_queries.argtypes = [vp, u64, u64, u32, vp, vp, vp, vp, vp]
        L.wl_gen_bench_docs.argtypes = [u64, u64, u64, vp, vp, vp, vp, vp]
        L.wl_gen_removed.restype = u64
        L.wl_gen_removed.argtypes = [u64, u64, C.c_double, vp]
        _lib = L
    return _lib


CORPUS_SEED = 0x5EEDC0DE
QUERY_SEED = 0x9E3779B9


@dataclass(frozen=True)
class Config:
    name: str
    cfg: int
    n_docs: int
    n_fields: int
    vocab: int
    len_min: Tuple[int, ...]
    len_max: Tuple[int, ...]
    n_queries: int
    query_mode: int          # 0 full term, 1 multi-term prefixes, 2 short prefix
    scorer: str              # "bm25" | "zero_to_one"
    boosts: Tuple[float, ...]
    removed_fraction: float = 0.0
    bench_shape: bool = False


CONFIGS = {
    # cfg 0 (reference CPU): 50k docs, 1 field; Zipf variant V = 2^16, 2 tokens/doc
    "cfg0": Config("cfg0", 0, 50_000, 1, 1 << 16, (2,), (2,), 1_000, 0, "bm25", (1.0,)),
    # cfg 0 bench shape: two random 5-letter words (benches/test_benchmark.rs), 1-2 char prefix queries
    "cfg0_bench": Config("cfg0_bench", 0, 50_000, 1, 1 << 16, (2,), (2,), 1_000, 2, "bm25", (1.0,), bench_shape=True),
    # cfg 1 (1xB200): 1M docs, 2 fields Zipf, 100k single-term BM25 queries
    "cfg1": Config("cfg1", 1, 1_000_000, 2, 1 << 18, (1, 8), (8, 40), 100_000, 0, "bm25", (1.0, 1.0)),
    # cfg 2 (1xB200): same corpus, 100k multi-term prefix queries, zero-to-one
    "cfg2": Config("cfg2", 1, 1_000_000, 2, 1 << 18, (1, 8), (8, 40), 100_000, 1, "zero_to_one", (1.0, 1.0)),
    # cfg 3 (8xB200): 10M docs, V = 2^20, 1M BM25 queries sharded by query
    "cfg3": Config("cfg3", 3, 10_000_000, 2, 1 << 20, (1, 8), (8, 40), 1_000_000, 0, "bm25", (1.0, 1.0)),
    # cfg 4 (8xB200): cfg 3 corpus, boosts [2.0, 0.5], 5% removed (pre-vacuum)
    "cfg4": Config("cfg4", 3, 10_000_000, 2, 1 << 20, (1, 8), (8, 40), 1_000_000, 0, "bm25", (2.0, 0.5), 0.05),
}


class Workload:
    def __init__(self, cfg: Config, n_docs: int | None = None, vocab: int | None = None):
        self.cfg = cfg
        self.n_docs = int(n_docs if n_docs is not None else cfg.n_docs)
        self.vocab = int(vocab if vocab is not None else cfg.vocab)
        self.seed = CORPUS_SEED + cfg.cfg
        lmin = np.asarray(cfg.len_min, dtype=np.uint32)
        lmax = np.asarray(cfg.len_max, dtype=np.uint32)
        self._h = lib().wl_new(self.seed, self.vocab, cfg.n_fields, lmin.ctypes.data, lmax.ctypes.data)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().wl_free(self._h)
            self._h = None

    def vocab_word(self, rank: int) -> str:
        buf = np.zeros(16, dtype=np.uint8)
        n = lib().wl_vocab_word(self._h, rank, buf.ctypes.data, 16)
        return bytes(buf[:n]).decode()

    def doc_chunks(self, chunk: int = 100_000) -> Iterator[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]:
        """Yields (keys, tok_bytes, tok_off, field_tok_count) for consecutive doc ranges; key = ordinal."""
        L = lib()
        F = self.cfg.n_fields
        for d0 in range(0, self.n_docs, chunk):
            d1 = min(self.n_docs, d0 + chunk)
            mt, mb = C.c_uint64(0), C.c_uint64(0)
            L.wl_doc_bounds(self._h, d1 - d0, C.byref(mt), C.byref(mb))
            tb = np.zeros(mb.value + 16, dtype=np.uint8)
            to = np.zeros(mt.value + 1, dtype=np.uint64)
            fc = np.zeros((d1 - d0) * F, dtype=np.uint32)
            nt, nb = C.c_uint64(0), C.c_uint64(0)
            if self.cfg.bench_shape:
                L.wl_gen_bench_docs(self.seed, d0, d1, tb.ctypes.data, to.ctypes.data, fc.ctypes.data, C.byref(nt), C.byref(nb))
            else:
                L.wl_gen_docs(self._h, d0, d1, tb.ctypes.data, to.ctypes.data, fc.ctypes.data, C.byref(nt), C.byref(nb))
            yield np.arange(d0, d1, dtype=np.uint64), tb[: nb.value + 1], to[: nt.value + 1], fc

    def queries(self, n_queries: int | None = None, mode: int | None = None) -> FlatQueries:
        n = int(n_queries if n_queries is not None else self.cfg.n_queries)
        mode = self.cfg.query_mode if mode is None else mode
        qoff = np.zeros(n + 1, dtype=np.uint64)
        toff = np.zeros(n * 4 + 1, dtype=np.uint64)
        tb = np.zeros(n * 4 * 10 + 16, dtype=np.uint8)
        nt, nb = C.c_uint64(0), C.c_uint64(0)
        lib().wl_gen_queries(self._h, QUERY_SEED + self.cfg.cfg, n, mode, qoff.ctypes.data, toff.ctypes.data,
                             tb.ctypes.data, C.byref(nt), C.byref(nb))
        return FlatQueries(qoff, toff[: nt.value + 1], tb[: nb.value + 1])

    def removed_ordinals(self) -> np.ndarray:
        if self.cfg.removed_fraction <= 0:
            return np.zeros(0, dtype=np.uint64)
        out = np.zeros(self.n_docs, dtype=np.uint64)
        n = lib().wl_gen_removed(self.seed, self.n_docs, self.cfg.removed_fraction, out.ctypes.data)
        return out[:n].copy()

    def build_into(self, index, chunk: int = 100_000) -> None:
        """Adds the whole corpus to an object with `add_documents_flat` (product Index or oracle)."""
        for keys, tb, to, fc in self.doc_chunks(chunk):
            index.add_documents_flat(keys, tb, to, fc)
