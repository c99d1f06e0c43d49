"""Builds the product library IN-TREE: probly_search_b200/_lib/libprobly_b200.so (host index
builder + CUDA engine, sm_100a) and libprobly_workload.so (synthetic corpus / query generator
used by tests and bench).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT, "libprobly_b200.so")
WLIB = os.path.join(OUT, "libprobly_workload.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-cudart", "shared", "-diag-suppress", "63,177",
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_variant(name: str, defs) -> str:
    """Tuning helper: the same library with extra -D flags, as _lib/libprobly_b200_<name>.so."""
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, f"libprobly_b200_{name}.so")
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "builder.cpp", "common.cpp", "image_io.cpp")]
    subprocess.check_call(["nvcc"] + NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-o", out] + srcs)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "builder.cpp", "common.cpp", "image_io.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("kernels.cuh", "common.hpp")] + [
        os.path.join(HERE, "..", "include", "probly_b200.h")]
    if force or _newer(LIB, deps):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
        subprocess.check_call(cmd)
    wsrc = os.path.join(CSRC, "workload.cpp")
    if os.path.exists(wsrc) and (force or _newer(WLIB, [wsrc])):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-o", WLIB, wsrc])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
