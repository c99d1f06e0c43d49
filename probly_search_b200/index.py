"""Host-side mirror of the reference's public surface for the query path.

`Index` keeps the names, argument meaning and error behaviour of `probly_search::Index`
(src/index.rs:35-199, src/query.rs:17-106): `Index(fields_num)`, `add_document(field_accessors,
tokenizer, key, doc)`, `remove_document(key)`, `vacuum()`, `query(query, score_calculator,
tokenizer, fields_boost) -> [QueryResult]`, plus the batch entry `query_batch`.  All work goes
through the C ABI (include/probly_b200.h); there is no Python implementation of the path.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Any, Callable, Hashable, List, Optional, Sequence

import numpy as np

from . import capi
from .score import scorer_params


@dataclass(frozen=True)
class QueryResult:
    """src/query.rs:9-15"""
    key: Any
    score: float


Tokenizer = Callable[[str], List[str]]          # src/lib.rs:14
FieldAccessor = Callable[[Any], List[str]]      # src/lib.rs:11


def _flat_tokens(tokens: Sequence[str]):
    enc = [t.encode("utf-8") for t in tokens]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    buf = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
    return buf, off


class FlatQueries:
    """A query batch in the layout of `pb_query_batch_desc`: every query already tokenized."""

    def __init__(self, query_term_off: np.ndarray, term_byte_off: np.ndarray, term_bytes: np.ndarray):
        self.query_term_off = np.ascontiguousarray(query_term_off, dtype=np.uint64)
        self.term_byte_off = np.ascontiguousarray(term_byte_off, dtype=np.uint64)
        self.term_bytes = np.ascontiguousarray(term_bytes, dtype=np.uint8)

    @property
    def n_queries(self) -> int:
        return len(self.query_term_off) - 1

    @classmethod
    def from_strings(cls, queries: Sequence[str], tokenizer: Tokenizer) -> "FlatQueries":
        toks: List[str] = []
        qoff = [0]
        for q in queries:
            toks.extend(tokenizer(q))
            qoff.append(len(toks))
        buf, off = _flat_tokens(toks)
        return cls(np.asarray(qoff, dtype=np.uint64), off, buf)

    def slice(self, lo: int, hi: int) -> "FlatQueries":
        t0, t1 = int(self.query_term_off[lo]), int(self.query_term_off[hi])
        b0, b1 = int(self.term_byte_off[t0]), int(self.term_byte_off[t1])
        return FlatQueries(self.query_term_off[lo:hi + 1] - np.uint64(t0),
                           self.term_byte_off[t0:t1 + 1] - np.uint64(b0),
                           np.concatenate([self.term_bytes[b0:b1], np.zeros(1, np.uint8)]))

    def terms_of(self, q: int) -> List[str]:
        out = []
        for t in range(int(self.query_term_off[q]), int(self.query_term_off[q + 1])):
            out.append(bytes(self.term_bytes[int(self.term_byte_off[t]):int(self.term_byte_off[t + 1])]).decode())
        return out


class BatchResults:
    """Per-query outputs of `pb_query_batch` (include/probly_b200.h `pb_query_results`)."""

    def __init__(self, n: int, k: int):
        kk = max(k, 1)
        self.k = k
        self.n_results = np.zeros(n, dtype=np.uint64)
        self.doc_digest = np.zeros(n, dtype=np.uint64)
        self.score_digest = np.zeros(n, dtype=np.uint64)
        self.topk_n = np.zeros(n, dtype=np.uint32)
        self.topk_doc = np.zeros((n, kk), dtype=np.uint32)
        self.topk_score = np.zeros((n, kk), dtype=np.float64)

    def c_struct(self) -> capi.QueryResults:
        return capi.QueryResults(self.n_results.ctypes.data, self.doc_digest.ctypes.data,
                                 self.score_digest.ctypes.data, self.topk_n.ctypes.data,
                                 self.topk_doc.ctypes.data, self.topk_score.ctypes.data)


class Index:
    """Mirror of `Index<T>` (src/index.rs:19-33).  Keys may be any hashable; the device works on
    dense ordinals and the host keeps ordinal -> key."""

    def __init__(self, fields_num: int, device: int = 0):
        self._L = capi.lib()
        self.fields_num = fields_num
        self.device = device
        h = C.c_void_p()
        capi.check(self._L.pb_builder_create(fields_num, C.byref(h)))
        self._b = h
        self._ix: Optional[C.c_void_p] = None
        self._image_dirty = True        # structure changed: re-flatten + upload
        self._live_dirty = False        # only the removed set / stats changed
        self._delta_dirty = False       # documents were added behind the resident image: delta segment
        self._ix_delta: Optional[C.c_void_p] = None
        self._n_main_docs = 0           # doc ordinals / posting rows the resident MAIN image covers
        self._n_main_rows = 0
        self._sid_main: Optional[np.ndarray] = None     # builder term id of every term ordinal (matches terms across segments)
        self._sid_delta: Optional[np.ndarray] = None
        self._key_to_id: dict = {}
        self._id_to_key: list = []
        self._ord_to_id: Optional[np.ndarray] = None
        self._batches = weakref.WeakSet()   # staged DeviceBatch objects: invalidated when the pb_index is recreated

    # -- on-disk image (the reference has no serialisation) -----------------------
    def save_image(self, path: str) -> None:
        """Writes the flattened image (what pb_index_create uploads) to `path` (csrc/image_io.cpp)."""
        im = self.flatten()             # always the FULL image (delta segments exist on the device only)
        capi.check(self._L.pb_image_save(C.byref(im), os.fsencode(path)))

    @classmethod
    def load_image(cls, path: str, device: int = 0) -> "Index":
        """A query-only Index served straight from an image file: no host builder exists, mutation
        raises, result keys are the u64 key ids stored in the image."""
        self = cls.__new__(cls)
        self._L = capi.lib()
        h = C.c_void_p()
        capi.check(self._L.pb_image_load(os.fsencode(path), C.byref(h)))
        self._image_file = h
        self._b = None
        self._ix = None
        self.device = device
        im = C.cast(self._L.pb_image_file_image(h), C.POINTER(capi.IndexImage)).contents
        self.fields_num = int(im.num_fields)
        self._image_dirty, self._live_dirty, self._delta_dirty = True, False, False
        self._ix_delta, self._n_main_docs, self._n_main_rows, self._sid_main, self._sid_delta = None, 0, 0, None, None
        self._key_to_id, self._id_to_key, self._ord_to_id = {}, [], None
        self._flat_keys = True
        self._batches = weakref.WeakSet()
        return self

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_image_file", None):
            self._L.pb_image_file_free(self._image_file)
            self._image_file = None
        if getattr(self, "_ix", None):
            self._drop_device_index()
        if getattr(self, "_b", None):
            self._L.pb_builder_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mutation (src/index.rs:77-199) --------------------------------------------------------
    def _key_id(self, key: Hashable) -> int:
        i = self._key_to_id.get(key)
        if i is None:
            i = len(self._id_to_key)
            self._key_to_id[key] = i
            self._id_to_key.append(key)
        return i

    def add_document(self, field_accessors: Sequence[FieldAccessor], tokenizer: Tokenizer, key: Hashable,
                     doc: Any) -> None:
        """src/index.rs:77-158"""
        toks: List[str] = []
        vcount: List[int] = []
        fcount: List[int] = []
        for i in range(self.fields_num):
            values = field_accessors[i](doc)
            fcount.append(len(values))
            for v in values:
                t = tokenizer(v)
                vcount.append(len(t))
                toks.extend(t)
        buf, off = _flat_tokens(toks)
        vc = np.asarray(vcount + [0], dtype=np.uint32)
        fc = np.asarray(fcount, dtype=np.uint32)
        d = capi.DocTokens(buf.ctypes.data, off.ctypes.data, vc.ctypes.data, fc.ctypes.data)
        capi.check(self._L.pb_builder_add_document(self._require_builder(), self._key_id(key), C.byref(d)))
        self._mark_added()

    def add_documents_flat(self, keys: np.ndarray, tok_bytes: np.ndarray, tok_off: np.ndarray,
                           field_tok_count: np.ndarray) -> None:
        """Bulk add with integer keys (pb_builder_add_documents): one value per field."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        if self._id_to_key:
            raise ValueError("add_documents_flat cannot be mixed with add_document on one index")
        self._flat_keys = True
        capi.check(self._L.pb_builder_add_documents(self._require_builder(), len(keys), keys.ctypes.data, tok_bytes.ctypes.data,
                                                    tok_off.ctypes.data, field_tok_count.ctypes.data))
        self._mark_added()

    def _mark_added(self) -> None:
        """Documents added behind a resident image become a small DELTA segment at the next query (SURVEY §8f-1:
        add_document stays cheap, src/index.rs:77-158); without a resident image the next sync flattens everything."""
        if self._ix is None or self._image_dirty:
            self._image_dirty = True
        else:
            self._delta_dirty = True

    def remove_document(self, key: Hashable) -> None:
        """src/index.rs:161-191 — lazy: the postings stay until vacuum()."""
        kid = key if getattr(self, "_flat_keys", False) else self._key_to_id.get(key)
        if kid is None:
            return
        capi.check(self._L.pb_builder_remove_document(self._require_builder(), int(kid)))
        self._live_dirty = True

    def vacuum(self) -> None:
        """src/index.rs:194-199"""
        capi.check(self._L.pb_builder_vacuum(self._require_builder()))
        self._image_dirty = True

    def info(self) -> capi.BuilderInfo:
        out = capi.BuilderInfo()
        capi.check(self._L.pb_builder_get_info(self._require_builder(), C.byref(out)))
        return out

    def flatten(self) -> capi.IndexImage:
        if getattr(self, "_image_file", None):
            return C.cast(self._L.pb_image_file_image(self._image_file), C.POINTER(capi.IndexImage)).contents
        im = capi.IndexImage()
        capi.check(self._L.pb_builder_flatten(self._require_builder(), C.byref(im)))
        return im

    def _require_builder(self):
        if self._b is None:
            raise capi.ProblyError(capi.PB_ERR_UNSUPPORTED, "this Index was loaded from an image file and cannot be mutated")
        return self._b

    # -- device image --------------------------------------------------------------------------
    # a delta segment is folded into the main image once it holds more than this share of the main image's rows
    DELTA_MAX_FRACTION = 0.25
    DELTA_MIN_ROWS = 1 << 16

    def _structure(self, from_doc: int) -> capi.IndexImage:
        """The image without posting columns (pb_builder_flatten_structure): node / term arrays, doc keys, live state."""
        im = capi.IndexImage()
        capi.check(self._L.pb_builder_flatten_structure(self._require_builder(), int(from_doc), C.byref(im)))
        return im

    def _term_ids_of_last_flatten(self) -> np.ndarray:
        n = C.c_uint64(0)
        self._L.pb_builder_flatten_term_ids(self._b, None, 0, C.byref(n))
        out = np.zeros(max(int(n.value), 1), dtype=np.uint32)
        capi.check(self._L.pb_builder_flatten_term_ids(self._b, out.ctypes.data, len(out), C.byref(n)))
        return out[: int(n.value)]

    def _apply_live_state(self, im) -> None:
        """Removed set / N / averages of the host index on the resident image.  With a delta segment attached the
        library splits the removed ordinals over the two segments and lets them exchange their per-term live counts
        (BM25's document frequency is over the whole index)."""
        nd = int(im.n_docs)
        words = np.ctypeslib.as_array(im.removed_bitmap, shape=((nd + 31) // 32 + 1,))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:nd]
        ords = np.ascontiguousarray(np.nonzero(bits)[0], dtype=np.uint32)
        avg = (C.c_double * 4)(*[im.field_avg[i] for i in range(4)])
        capi.check(self._L.pb_index_set_live_state(self._ix, ords.ctypes.data, len(ords), im.n_live_docs, avg))
        if self._ord_to_id is None or len(self._ord_to_id) != nd:      # ordinals -> keys only change when docs are added / vacuumed
            self._ord_to_id = np.ctypeslib.as_array(im.doc_key, shape=(nd,)).copy() if nd else np.zeros(0, np.uint64)

    def _drop_delta(self) -> None:
        if self._ix_delta is not None:
            if self._ix is not None:                   # the main image owns its delta segment: detaching destroys it
                capi.check(self._L.pb_index_attach_delta(self._ix, None, None, 0, None, 0))
            self._ix_delta = None
            self._sid_delta = None

    def compact(self) -> None:
        """Folds the delta segment into one image (the GPU analogue of a merge; also what vacuum() forces)."""
        if self._ix_delta is not None or self._delta_dirty:
            self._image_dirty = True
        self.sync_device()

    @property
    def n_segments(self) -> int:
        return (1 if self._ix is not None else 0) + (1 if self._ix_delta is not None else 0)

    def sync_device(self) -> None:
        """Brings the HBM image up to date with the host index: a full flatten + upload the first time and after
        vacuum(); a small DELTA segment (the rows of the documents added since, under the current trie) after
        add_document; just the removed mask / N / avg / idf after remove_document."""
        if self._ix is not None and not self._image_dirty and not self._live_dirty and not self._delta_dirty:
            return
        if getattr(self, "_image_file", None):          # served from a file: one immutable segment
            im = self.flatten()
            if self._ix is None:
                h = C.c_void_p()
                capi.check(self._L.pb_index_create(C.byref(im), self.device, C.byref(h)))
                self._ix = h
                nd = int(im.n_docs)
                self._ord_to_id = np.ctypeslib.as_array(im.doc_key, shape=(nd,)).copy() if nd else np.zeros(0, np.uint64)
            self._image_dirty = self._live_dirty = self._delta_dirty = False
            return
        if self._ix is not None and not self._image_dirty and self._delta_dirty:
            info = self.info()
            delta_rows = int(info.n_rows) - self._n_main_rows
            if delta_rows > max(self.DELTA_MIN_ROWS, self.DELTA_MAX_FRACTION * self._n_main_rows):
                self._image_dirty = True                  # the delta has grown: fold it into the main image
        if self._ix is None or self._image_dirty:
            self._drop_delta()
            if self._ix is not None:
                self._drop_device_index()
            # the posting columns are flattened on the device from the builder's append log; the host flattens the
            # small structures only (trie, term table, doc keys, live state)
            h = C.c_void_p()
            capi.check(self._L.pb_index_create_from_builder(self._require_builder(), 0, self.device, C.byref(h)))
            im = self._structure(0)
            self._ix = h
            self._n_main_docs, self._n_main_rows = int(im.n_docs), int(im.n_rows)
            self._sid_main = self._term_ids_of_last_flatten()
            self._apply_live_state(im)
        elif self._delta_dirty:
            h = C.c_void_p()
            capi.check(self._L.pb_index_create_from_builder(self._require_builder(), self._n_main_docs, self.device, C.byref(h)))
            im = self._structure(self._n_main_docs)
            self._sid_delta = self._term_ids_of_last_flatten()
            # pb_index_attach_delta: the main image takes ownership (a previous delta is destroyed), the two segments
            # exchange their per-term live counts, and the query entry points answer for both from now on
            rc = self._L.pb_index_attach_delta(self._ix, h, self._sid_main.ctypes.data, len(self._sid_main),
                                               self._sid_delta.ctypes.data, len(self._sid_delta))
            if rc != capi.PB_OK:
                self._L.pb_index_destroy(h)
                capi.check(rc)
            self._ix_delta = h
            self._apply_live_state(im)
        else:                                             # only remove_document happened
            im = self._structure(self._n_main_docs if self._ix_delta is not None else 0)
            self._apply_live_state(im)
        self._image_dirty = self._live_dirty = self._delta_dirty = False

    def set_live_state(self, removed_ordinals: np.ndarray, n_live_docs: int, field_avg: Sequence[float]) -> None:
        """pb_index_set_live_state on the resident image: the FULL removed set, the live doc count and the
        per-field averages of the pre-vacuum state (src/index.rs:161-191) — no re-flatten, no re-upload.
        This is how an index served from an image file (no host builder) follows remove_document."""
        self.sync_device()
        ords = np.ascontiguousarray(removed_ordinals, dtype=np.uint32)
        avg = (C.c_double * 4)(*(list(field_avg) + [0.0] * 4)[:4])
        capi.check(self._L.pb_index_set_live_state(self._ix, ords.ctypes.data, len(ords), int(n_live_docs), avg))

    def live_state(self):
        """(removed ordinals, live doc count, field averages) of the host index, as set_live_state takes them."""
        im = self.flatten()
        nd = int(im.n_docs)
        words = np.ctypeslib.as_array(im.removed_bitmap, shape=((nd + 31) // 32 + 1,))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:nd]
        ords = np.ascontiguousarray(np.nonzero(bits)[0], dtype=np.uint32)
        return ords, int(im.n_live_docs), [float(im.field_avg[i]) for i in range(self.fields_num)]

    def _drop_device_index(self) -> None:
        """Destroys the pb_index; staged batches hold its raw handle, so they are closed first and any
        later use of them raises instead of touching freed memory."""
        for b in list(getattr(self, "_batches", ())):
            b._invalidate()
        self._L.pb_index_destroy(self._ix)      # destroys an attached delta segment with it
        self._ix = None
        self._ix_delta = None
        self._sid_delta = None

    def _key_of_ord(self, o: int):
        kid = int(self._ord_to_id[o])
        return kid if getattr(self, "_flat_keys", False) else self._id_to_key[kid]

    def _desc(self, fq: FlatQueries, score_calculator, fields_boost: Sequence[float], top_k: int):
        scorer, k1, b = scorer_params(score_calculator)
        boosts = np.asarray(list(fields_boost), dtype=np.float64)
        d = capi.QueryBatchDesc(fq.n_queries, fq.query_term_off.ctypes.data, fq.term_byte_off.ctypes.data,
                                fq.term_bytes.ctypes.data, scorer, k1, b, boosts.ctypes.data, len(boosts), top_k)
        return d, boosts

    # -- queries -------------------------------------------------------------------------------
    def expand_term(self, term: str) -> List[str]:
        """src/query.rs:109-126 (private there; exposed for the expansion-order goldens)."""
        self.compact()
        tb = np.frombuffer(term.encode("utf-8") + b"\0", dtype=np.uint8)
        n, need = C.c_uint64(0), C.c_uint64(0)
        capi.check(self._L.pb_index_expand_term(self._ix, tb.ctypes.data, len(tb) - 1, None, 0, C.byref(n), C.byref(need)))
        if n.value == 0:
            return []
        out = np.zeros(need.value + 1, dtype=np.uint8)
        capi.check(self._L.pb_index_expand_term(self._ix, tb.ctypes.data, len(tb) - 1, out.ctypes.data, need.value,
                                                C.byref(n), C.byref(need)))
        return bytes(out[: need.value]).decode("utf-8").split("\n")

    def query_full_flat(self, fq: FlatQueries, score_calculator, fields_boost: Sequence[float],
                        cap: Optional[int] = None):
        """Full result sets of every query of the batch: arrays (query, doc ordinal, score), unordered (with a delta
        segment attached the library concatenates the two segments' sets: a document lives in exactly one)."""
        self.sync_device()
        d, _keep = self._desc(fq, score_calculator, fields_boost, 0)
        cap = int(cap if cap is not None else max(1024, fq.n_queries * 64))
        while True:
            oq = np.zeros(cap, dtype=np.uint32)
            od = np.zeros(cap, dtype=np.uint32)
            os_ = np.zeros(cap, dtype=np.float64)
            n = C.c_uint64(0)
            rc = self._L.pb_query_full(self._ix, C.byref(d), cap, oq.ctypes.data, od.ctypes.data, os_.ctypes.data, C.byref(n))
            if rc == capi.PB_ERR_CAPACITY:
                cap = int(n.value) + 16
                continue
            capi.check(rc)
            return oq[: n.value], od[: n.value], os_[: n.value]

    def query(self, query: str, score_calculator, tokenizer: Tokenizer, fields_boost: Sequence[float]) -> List[QueryResult]:
        """src/query.rs:21-106.  Result order: score descending; exactly tied scores by document
        ordinal ascending (the reference leaves ties in hash order, SURVEY §3.4 rule 10)."""
        fq = FlatQueries.from_strings([query], tokenizer)
        _, docs, scores = self.query_full_flat(fq, score_calculator, fields_boost)
        order = np.lexsort((docs, -scores))
        return [QueryResult(self._key_of_ord(int(docs[i])), float(scores[i])) for i in order]

    def query_batch_flat(self, fq: FlatQueries, score_calculator, fields_boost: Sequence[float], top_k: int = 10) -> BatchResults:
        """pb_query_batch: host buffers in, host buffers out (both segments answer when a delta is attached)."""
        self.sync_device()
        d, _keep = self._desc(fq, score_calculator, fields_boost, top_k)
        res = BatchResults(fq.n_queries, top_k)
        rs = res.c_struct()
        capi.check(self._L.pb_query_batch(self._ix, C.byref(d), C.byref(rs)))
        return res

    def query_batch(self, queries: Sequence[str], score_calculator, tokenizer: Tokenizer,
                    fields_boost: Sequence[float], top_k: int = 10) -> List[List[QueryResult]]:
        """Batch form of `query`: the top_k best results of every query."""
        fq = FlatQueries.from_strings(queries, tokenizer)
        r = self.query_batch_flat(fq, score_calculator, fields_boost, top_k)
        out = []
        for q in range(fq.n_queries):
            n = int(r.topk_n[q])
            out.append([QueryResult(self._key_of_ord(int(r.topk_doc[q, i])), float(r.topk_score[q, i])) for i in range(n)])
        return out

    def last_stats(self) -> dict:
        s = capi.BatchStats()
        capi.check(self._L.pb_index_last_stats(self._ix, C.byref(s)))
        return s.as_dict()

    def device_layout(self) -> dict:
        """How the posting columns are held in HBM (pb_device_layout): narrow u16 codes or u32 columns."""
        self.compact()
        d = capi.DeviceLayout()
        capi.check(self._L.pb_index_device_layout(self._ix, C.byref(d)))
        return {"narrow": bool(d.narrow), "bytes_per_row": int(d.bytes_per_row), "posting_bytes": int(d.posting_bytes),
                "fl_bits": [int(x) for x in d.fl_bits][: self.fields_num]}

    def term_df_live(self) -> np.ndarray:
        self.compact()
        im = self.flatten()
        out = np.zeros(max(int(im.n_terms), 1), dtype=np.uint64)
        capi.check(self._L.pb_index_term_df_live(self._ix, out.ctypes.data, len(out)))
        return out[: int(im.n_terms)]


class DeviceBatch:
    """Staged form of a batch (pb_batch_create / run / fetch): upload once, run many times with
    the inputs resident in HBM.  With `set_gather(comm, slot)` every run ends with ONE
    ncclAllGather of the packed result block on the batch's stream (multi-GPU)."""

    def __init__(self, index: Index, fq: FlatQueries, score_calculator, fields_boost: Sequence[float], top_k: int = 10):
        index.compact()                  # a staged batch runs on ONE image: a pending delta segment is folded in first
        self._L = index._L
        self.index = index
        self.fq = fq
        self.top_k = top_k
        self._calc = score_calculator
        self._fields_boost = list(fields_boost)
        d, self._boosts = index._desc(fq, score_calculator, fields_boost, top_k)
        h = C.c_void_p()
        capi.check(self._L.pb_batch_create(index._ix, C.byref(d), C.byref(h)))
        self._h = h
        self._stale = False
        self._comm = None
        index._batches.add(self)

    def _handle(self):
        if self._stale or not getattr(self, "_h", None):
            raise capi.ProblyError(capi.PB_ERR_INVALID, "this DeviceBatch was staged on a device image that has since been "
                                   "rebuilt (add_document / vacuum) or closed; stage a new one")
        return self._h

    def _invalidate(self) -> None:
        self.close()
        self._stale = True

    def reload(self, fq: FlatQueries) -> None:
        """Stages a new query batch into the same pb_batch (keeps stream, workspace and gather)."""
        d, self._boosts = self.index._desc(fq, self._calc, self._fields_boost, self.top_k)
        capi.check(self._L.pb_batch_reload(self._handle(), C.byref(d)))
        self.fq = fq

    def set_gather(self, comm, slot_queries: int) -> None:
        capi.check(self._L.pb_batch_set_gather(self._handle(), comm._h if comm is not None else None, int(slot_queries)))
        self._comm = comm
        self._slot = int(slot_queries)

    def _fresh(self):
        """A staged batch answers from the image it was staged on: removals follow (live state only), added
        documents need a new batch (the delta segment is a second image)."""
        ix = self.index
        if ix._image_dirty or ix._delta_dirty or ix._ix_delta is not None:
            raise capi.ProblyError(capi.PB_ERR_INVALID, "documents were added to the index after this DeviceBatch was staged; "
                                   "stage a new one (Index.compact() folds the delta segment in)")
        if ix._live_dirty:
            ix.sync_device()
        return self._handle()

    def run(self) -> None:
        capi.check(self._L.pb_batch_run(self._fresh()))

    def run_local(self) -> None:
        capi.check(self._L.pb_batch_run_local(self._fresh()))

    def fetch(self) -> BatchResults:
        res = BatchResults(self.fq.n_queries, self.top_k)
        rs = res.c_struct()
        capi.check(self._L.pb_batch_fetch(self._handle(), C.byref(rs)))
        return res

    def fetch_gathered(self, n_total: Optional[int] = None) -> BatchResults:
        """Results of ALL ranks in global query order (rank r's query i = r * slot + i)."""
        if self._comm is None:
            raise capi.ProblyError(capi.PB_ERR_INVALID, "no gather attached")
        n = int(n_total if n_total is not None else self._comm.world * self._slot)
        res = BatchResults(n, self.top_k)
        rs = res.c_struct()
        capi.check(self._L.pb_batch_fetch_gathered(self._handle(), n, C.byref(rs)))
        return res

    def stats(self) -> dict:
        s = capi.BatchStats()
        capi.check(self._L.pb_batch_get_stats(self._handle(), C.byref(s)))
        return s.as_dict()

    def device_results(self) -> capi.QueryResults:
        """Device pointers of the result buffers."""
        r = capi.QueryResults()
        capi.check(self._L.pb_batch_device_results(self._handle(), C.byref(r)))
        return r

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.pb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
