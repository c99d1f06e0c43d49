"""CPU-only: the C-ABI library loads and exports every symbol include/probly_b200.h declares, and
the query entry points fail loudly (PB_ERR_NO_DEVICE) instead of falling back when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from probly_search_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "probly_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/probly_b200.h but not exported"
    assert sorted(capi.EXPORTS) == syms


def test_struct_sizes_match_header():
    # compile a tiny C program against the header and compare sizeof with the ctypes mirrors
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include "probly_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(pb_doc_tokens), sizeof(pb_builder_info), sizeof(pb_index_image),
             sizeof(pb_query_batch_desc), sizeof(pb_query_results), sizeof(pb_batch_stats), sizeof(pb_device_layout));
      return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [capi.DocTokens, capi.BuilderInfo, capi.IndexImage, capi.QueryBatchDesc, capi.QueryResults, capi.BatchStats,
               capi.DeviceLayout]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_cpu_fallback_without_device():
    L = capi.lib()
    if L.pb_device_count() > 0:
        pytest.skip("a GPU is present")
    from probly_search_b200 import Index, score
    ix = Index(1)
    ix.add_document([lambda d: [d]], lambda s: s.split(" "), 0, "a b")
    with pytest.raises(capi.ProblyError) as e:
        ix.query("a", score.bm25.new(), lambda s: s.split(" "), [1.0])
    assert e.value.code == capi.PB_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "probly_search_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f in ("workload.py", "workload.cpp", "index.py"), f


def test_new_entry_points_reject_bad_arguments_without_a_device():
    L = capi.lib()
    assert L.pb_index_device_layout(None, None) == capi.PB_ERR_INVALID
    g = C.c_double(0.0)
    assert L.pb_device_read_bandwidth(0, 16, 1, C.byref(g)) == capi.PB_ERR_INVALID          # buffer too small
    if L.pb_device_count() == 0:
        assert L.pb_device_read_bandwidth(0, 1 << 20, 1, C.byref(g)) == capi.PB_ERR_NO_DEVICE
    assert L.pb_image_save(None, b"/tmp/x") == capi.PB_ERR_INVALID
    h = C.c_void_p()
    assert L.pb_image_load(None, C.byref(h)) == capi.PB_ERR_INVALID
    assert L.pb_image_file_image(None) is None
    L.pb_image_file_free(None)                                                               # no-op, must not crash
