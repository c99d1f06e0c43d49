"""-m gpu: compiles tests/cpp/integration_test.cpp (the reference's integration tests restated on
the header-only C++ mirror include/probly_b200.hpp) against the in-tree C-ABI library and runs it."""
import os
import subprocess

import pytest

from probly_search_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    capi.lib()
    exe = str(tmp_path / "integration_test")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "integration_test.cpp"), "-o", exe,
                           "-L", libdir, "-lprobly_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_cpp_mirror_compiles(tmp_path):
    _compile(tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_integration(tmp_path):
    exe = _compile(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "OK" in out.stdout
