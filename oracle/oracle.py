"""ctypes wrapper around oracle/probly_oracle.cpp — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (probly_search_b200) never does.

The wrapper mirrors the reference's surface (`Index::new`, `add_document`,
`remove_document`, `vacuum`, `query`, `expand_term`; src/index.rs, src/query.rs) with
the test tokenizer of src/lib.rs:42-44 (split on a single ' ', empty tokens kept).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "probly_oracle.cpp")
_SO = os.path.join(_HERE, "_build", "liboracle.so")

BM25 = 0
ZERO_TO_ONE = 1


def build(force: bool = False) -> str:
    """g++ the oracle into oracle/_build/liboracle.so (no FMA contraction, so f64 matches rustc)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["g++", "-O3", "-march=native", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
             "-pthread", "-o", _SO, _SRC])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_SO):
                raise
        L = C.CDLL(_SO)
        vp, u64, u32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_double
        P = C.POINTER
        L.orc_index_new.restype = vp
        L.orc_index_new.argtypes = [u32]
        L.orc_index_free.argtypes = [vp]
        L.orc_add_document.argtypes = [vp, u64, vp, vp, vp, vp]
        L.orc_add_documents.argtypes = [vp, u64, vp, vp, vp, vp]
        L.orc_remove_document.argtypes = [vp, u64]
        L.orc_vacuum.argtypes = [vp]
        for f in ("orc_docs_len", "orc_arena_doc_len", "orc_arena_index_len", "orc_score_calls", "orc_count_nodes"):
            getattr(L, f).restype = u64
            getattr(L, f).argtypes = [vp]
        L.orc_field_stats.argtypes = [vp, vp, vp]
        L.orc_children_chars.restype = u64
        L.orc_children_chars.argtypes = [vp, vp, u64, vp, u64]
        L.orc_postings.restype = u64
        L.orc_postings.argtypes = [vp, vp, u64, vp, vp, u64]
        L.orc_expand_term.restype = u64
        L.orc_expand_term.argtypes = [vp, vp, u64, vp, u64, vp]
        L.orc_query.restype = u64
        L.orc_query.argtypes = [vp, vp, vp, u64, C.c_int, dbl, dbl, vp, vp, vp, u64]
        L.orc_query_batch.restype = dbl
        L.orc_query_batch.argtypes = [vp, u64, vp, vp, vp, C.c_int, dbl, dbl, vp, u32, u32,
                                      vp, vp, vp, vp, vp, vp, vp]
        L.orc_doc_hash.restype = u64
        L.orc_doc_hash.argtypes = [u64]
        L.orc_score_hash.restype = u64
        L.orc_score_hash.argtypes = [u64, dbl]
        _lib = L
    return _lib


def tokenizer(s: str) -> List[str]:
    """src/lib.rs:42-44 — `s.split(' ')`; keeps empty tokens."""
    return s.split(" ")


def _flat(tokens: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    enc = [t.encode("utf-8") for t in tokens]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    buf = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
    return buf, off


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


class OracleIndex:
    """Mirror of `Index<usize>` (src/index.rs:19-33) backed by the C++ restatement."""

    def __init__(self, fields_num: int):
        self.fields_num = fields_num
        self._h = lib().orc_index_new(fields_num)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_index_free(self._h)
            self._h = None

    # src/index.rs:77-158.  `field_values[i]` = list of raw string values of field i.
    def add_document(self, key: int, field_values: Sequence[Sequence[str]], tok=tokenizer) -> None:
        toks: List[str] = []
        vcount: List[int] = []
        fcount: List[int] = []
        for vals in field_values:
            fcount.append(len(vals))
            for v in vals:
                t = tok(v)
                vcount.append(len(t))
                toks.extend(t)
        buf, off = _flat(toks)
        vc = np.asarray(vcount + [0], dtype=np.uint32)
        fc = np.asarray(fcount, dtype=np.uint32)
        lib().orc_add_document(self._h, key, _p(buf), _p(off), _p(vc), _p(fc))

    def add_documents_flat(self, keys: np.ndarray, tok_bytes: np.ndarray, tok_off: np.ndarray,
                           field_tok_count: np.ndarray) -> None:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        lib().orc_add_documents(self._h, len(keys), _p(keys), _p(tok_bytes), _p(tok_off), _p(field_tok_count))

    def remove_document(self, key: int) -> None:
        lib().orc_remove_document(self._h, key)

    def vacuum(self) -> None:
        lib().orc_vacuum(self._h)

    def docs_len(self) -> int:
        return lib().orc_docs_len(self._h)

    def arena_doc_len(self) -> int:
        return lib().orc_arena_doc_len(self._h)

    def arena_index_len(self) -> int:
        return lib().orc_arena_index_len(self._h)

    def count_nodes(self) -> int:
        return lib().orc_count_nodes(self._h)

    def field_stats(self) -> List[Tuple[int, float]]:
        s = np.zeros(self.fields_num, dtype=np.uint64)
        a = np.zeros(self.fields_num, dtype=np.float64)
        lib().orc_field_stats(self._h, _p(s), _p(a))
        return [(int(x), float(y)) for x, y in zip(s, a)]

    def children_chars(self, path: str) -> List[str]:
        b = np.frombuffer(path.encode() + b"\0", dtype=np.uint8).copy()
        out = np.zeros(4096, dtype=np.uint32)
        n = lib().orc_children_chars(self._h, _p(b), len(path.encode()), _p(out), len(out))
        return [chr(c) for c in out[:n]]

    def postings(self, term: str) -> List[Tuple[int, List[int]]]:
        b = np.frombuffer(term.encode() + b"\0", dtype=np.uint8).copy()
        n = lib().orc_postings(self._h, _p(b), len(term.encode()), 0, 0, 0)
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        tf = np.zeros(max(n, 1) * self.fields_num, dtype=np.uint64)
        lib().orc_postings(self._h, _p(b), len(term.encode()), _p(keys), _p(tf), n)
        tf = tf.reshape(-1, self.fields_num)
        return [(int(keys[i]), [int(x) for x in tf[i]]) for i in range(n)]

    def expand_term(self, term: str) -> List[str]:
        b = np.frombuffer(term.encode() + b"\0", dtype=np.uint8).copy()
        need = C.c_uint64(0)
        n = lib().orc_expand_term(self._h, _p(b), len(term.encode()), 0, 0, C.byref(need))
        if n == 0:
            return []
        out = np.zeros(need.value + 1, dtype=np.uint8)
        lib().orc_expand_term(self._h, _p(b), len(term.encode()), _p(out), need.value, C.byref(need))
        return bytes(out[: need.value]).decode().split("\n")

    # src/query.rs:21-106 — returns [(key, score)] in the comparison order of
    # src/lib.rs:54-58 (score desc, key asc).
    def query(self, query: str, scorer: int = BM25, fields_boost: Iterable[float] | None = None,
              k1: float = 1.2, b: float = 0.75, tok=tokenizer) -> List[Tuple[int, float]]:
        return self.query_tokens(tok(query), scorer, fields_boost, k1, b)

    def query_tokens(self, tokens: Sequence[str], scorer: int = BM25,
                     fields_boost: Iterable[float] | None = None, k1: float = 1.2,
                     b: float = 0.75) -> List[Tuple[int, float]]:
        boosts = np.asarray(list(fields_boost) if fields_boost is not None else [1.0] * self.fields_num,
                            dtype=np.float64)
        buf, off = _flat(tokens)
        cap = max(int(self.docs_len()), 1)
        keys = np.zeros(cap, dtype=np.uint64)
        scores = np.zeros(cap, dtype=np.float64)
        n = lib().orc_query(self._h, _p(buf), _p(off), len(tokens), scorer, k1, b, _p(boosts),
                            _p(keys), _p(scores), cap)
        assert n <= cap
        return [(int(keys[i]), float(scores[i])) for i in range(n)]

    def query_batch_flat(self, query_tok_off: np.ndarray, tok_bytes: np.ndarray, tok_off: np.ndarray,
                         scorer: int, fields_boost: Sequence[float], top_k: int, k1: float = 1.2,
                         b: float = 0.75, n_threads: int = 1) -> dict:
        """Whole batch in C++ (timed).  Returns per-query count / digests / top-k and the elapsed seconds."""
        nq = len(query_tok_off) - 1
        boosts = np.asarray(list(fields_boost), dtype=np.float64)
        out = {
            "n_results": np.zeros(nq, dtype=np.uint64),
            "doc_digest": np.zeros(nq, dtype=np.uint64),
            "score_digest": np.zeros(nq, dtype=np.uint64),
            "topk_n": np.zeros(nq, dtype=np.uint32),
            "topk_key": np.zeros(nq * max(top_k, 1), dtype=np.uint64),
            "topk_score": np.zeros(nq * max(top_k, 1), dtype=np.float64),
        }
        calls = C.c_uint64(0)
        qo = np.ascontiguousarray(query_tok_off, dtype=np.uint64)
        secs = lib().orc_query_batch(self._h, nq, _p(qo), _p(tok_bytes), _p(tok_off), scorer, k1, b,
                                     _p(boosts), top_k, n_threads, _p(out["n_results"]),
                                     _p(out["doc_digest"]), _p(out["score_digest"]), _p(out["topk_n"]),
                                     _p(out["topk_key"]), _p(out["topk_score"]), C.byref(calls))
        out["seconds"] = secs
        out["score_calls"] = calls.value
        out["topk_key"] = out["topk_key"].reshape(nq, max(top_k, 1))
        out["topk_score"] = out["topk_score"].reshape(nq, max(top_k, 1))
        return out


def doc_hash(doc: int) -> int:
    return lib().orc_doc_hash(doc)


def score_hash(doc: int, score: float) -> int:
    return lib().orc_score_hash(doc, score)
