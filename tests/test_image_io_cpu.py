"""CPU-only: the on-disk image format (csrc/image_io.cpp) round-trips a flattened index bit for bit
and rejects truncated / corrupted / foreign files without aborting."""
import ctypes as C
import os

import numpy as np
import pytest

from probly_search_b200 import Index, capi
from tests import helpers as H


def _index():
    ix = Index(2)
    tok = lambda s: s.split(" ")
    docs = ["abc abd xyz", "ab abcde abc", "xyz ab abd abd", "abd abc q", "zz abcde", "hé llo wörld", ""]
    for k, d in enumerate(docs):
        ix.add_document([lambda d: [d], lambda d: [d[::-1], d]], tok, 100 + k, d)
    ix.remove_document(103)
    return ix


def test_round_trip_is_bit_exact(tmp_path):
    ix = _index()
    p = str(tmp_path / "ix.pbimg")
    ix.save_image(p)
    im0 = ix.flatten()
    a0 = H.image_arrays(im0)
    ld = Index.load_image(p)
    im1 = ld.flatten()
    for f in ("version", "num_fields", "n_nodes", "n_edges", "n_terms", "n_rows", "n_rows_padded", "n_docs",
              "max_term_bytes", "n_removed", "n_live_docs"):
        assert getattr(im0, f) == getattr(im1, f), f
    assert list(im0.max_tf) == list(im1.max_tf) and list(im0.max_fl) == list(im1.max_fl)
    assert [x for x in im0.field_avg] == [x for x in im1.field_avg]
    a1 = H.image_arrays(im1)
    for k, v in a0.items():
        if isinstance(v, list):
            for x, y in zip(v, a1[k]):
                np.testing.assert_array_equal(x, y)
        else:
            np.testing.assert_array_equal(v, a1[k])
    assert os.path.getsize(p) % 64 == 0
    with pytest.raises(capi.ProblyError):           # a loaded image is query-only
        ld.remove_document(100)
    ld.close()


def test_corrupt_and_foreign_files_are_rejected(tmp_path):
    ix = _index()
    p = str(tmp_path / "ix.pbimg")
    ix.save_image(p)
    raw = bytearray(open(p, "rb").read())
    L = capi.lib()

    def load(path):
        h = C.c_void_p()
        rc = L.pb_image_load(os.fsencode(path), C.byref(h))
        if rc == 0:
            L.pb_image_file_free(h)
        return rc

    assert load(p) == 0
    bad = str(tmp_path / "bad.pbimg")
    flipped = bytearray(raw); flipped[len(raw) - 100] ^= 0x40          # payload bit flip -> checksum
    open(bad, "wb").write(flipped)
    assert load(bad) == capi.PB_ERR_INVALID and "checksum" in capi.last_error()
    open(bad, "wb").write(raw[: len(raw) // 2])                          # truncated
    assert load(bad) == capi.PB_ERR_INVALID
    open(bad, "wb").write(b"not an image at all" * 10)                   # foreign
    assert load(bad) == capi.PB_ERR_INVALID and "magic" in capi.last_error()
    hdr = bytearray(raw); hdr[16 + 8] ^= 0x01                            # n_nodes changed -> section sizes disagree
    open(bad, "wb").write(hdr)
    assert load(bad) == capi.PB_ERR_INVALID
    assert load(str(tmp_path / "missing.pbimg")) == capi.PB_ERR_INVALID
