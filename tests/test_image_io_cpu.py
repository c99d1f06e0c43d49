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


def test_header_scalars_are_covered_by_the_checksum(tmp_path):
    """max_tf / max_fl / field_avg / n_live_docs size the device's u16 posting codes and the BM25 table: a flipped
    bit there must not load (round-1 advisory: only the section payloads were summed)."""
    ix = _index()
    p = str(tmp_path / "ix.pbimg")
    ix.save_image(p)
    raw = bytearray(open(p, "rb").read())
    L = capi.lib()
    # scalar block starts at byte 16: u32 version, u32 num_fields, 6 x u64, u32 max_term_bytes, u32 pad, max_tf[4] ...
    off_max_tf = 16 + 8 + 6 * 8 + 8
    for off in (off_max_tf, off_max_tf + 16, off_max_tf + 32 + 8, off_max_tf + 32 + 16):      # max_tf[0], max_fl[0], n_live_docs, field_avg[0]
        bad = bytearray(raw)
        bad[off] ^= 0x01
        q = str(tmp_path / "bad.pbimg")
        open(q, "wb").write(bad)
        h = C.c_void_p()
        assert L.pb_image_load(os.fsencode(q), C.byref(h)) == capi.PB_ERR_INVALID, off


def test_structurally_broken_images_are_refused():
    """The kernels index with the image's offsets and ordinals unchecked: validate_image (run by pb_image_load and
    by pb_index_create on any image) refuses out-of-range values instead of letting the device read out of bounds."""
    L = capi.lib()
    ix = _index()
    im = ix.flatten()
    p = "/tmp/pb_valid_probe.pbimg"

    def save_rc(image):
        return L.pb_image_save(C.byref(image), os.fsencode(p))

    assert save_rc(im) == 0
    h = C.c_void_p()
    assert L.pb_image_load(os.fsencode(p), C.byref(h)) == 0
    L.pb_image_file_free(h)

    def broken(mutate):
        """save a copy of the image with one array value changed, then try to load it"""
        a = H.image_arrays(im)
        keep = mutate(a)
        try:
            assert save_rc(im) == 0                  # saving does not validate the payload ...
            hh = C.c_void_p()
            rc = L.pb_image_load(os.fsencode(p), C.byref(hh))      # ... loading does
            if rc == 0:
                L.pb_image_file_free(hh)
            return rc
        finally:
            keep()

    def poke(arr, idx, val):
        def m(a):
            old = int(a[arr][idx])
            a[arr][idx] = val
            return lambda: a[arr].__setitem__(idx, old)
        return m

    def poke_block(word, val):                       # post_blocks is tile-blocked: word 0 = doc of row 0, word 128 = tf[0] of row 0
        def m(a):
            old = int(im.post_blocks[word])
            im.post_blocks[word] = val
            return lambda: im.post_blocks.__setitem__(word, old)
        return m

    nd, nn, nt = int(im.n_docs), int(im.n_nodes), int(im.n_terms)
    assert broken(poke_block(0, nd + 5)) == capi.PB_ERR_INVALID and "doc ordinals" in capi.last_error()
    assert broken(poke("edge_child", 0, nn)) == capi.PB_ERR_INVALID
    assert broken(poke("term_row_begin", 1, 10**9)) == capi.PB_ERR_INVALID
    assert broken(poke("node_term_hi", 0, nt + 1)) == capi.PB_ERR_INVALID
    assert broken(poke_block(128, int(im.max_tf[0]) + 1)) == capi.PB_ERR_INVALID and "max_tf" in capi.last_error()
    hh = C.c_void_p()
    assert save_rc(im) == 0 and L.pb_image_load(os.fsencode(p), C.byref(hh)) == 0      # everything restored: loads again
    L.pb_image_file_free(hh)
    os.unlink(p)
