"""Builds the product library IN-TREE: probly_search_b200/_lib/libprobly_b200.so (host index
builder + CUDA engine, sm_100a) and libprobly_workload.so (synthetic corpus / query generator
used by tests and bench).  nvcc cross-compiles without a GPU.

The per-field-count kernels (csrc/kernels_f.cu, -DPB_F=1..4) and the engine are separate
translation units compiled in parallel into _lib/obj/ and linked into one shared library."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT, "libprobly_b200.so")
WLIB = os.path.join(OUT, "libprobly_workload.so")

NVCC_COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-diag-suppress", "63,177",
]
LINK_FLAGS = ["-Wno-deprecated-gpu-targets", "-shared", "-cudart", "shared", "-Xlinker", "--no-as-needed", "-ldl"]

HEADERS = ["kernels.cuh", "plan_kernels.cuh", "union_kernels.cuh", "union_warp_kernel.cuh", "field_ops.hpp", "common.hpp", "group.hpp"]
# (object name, source, extra flags)
UNITS = [("engine", "engine.cu", []), ("group", "group.cu", [])] + \
        [(f"kernels_f{f}", "kernels_f.cu", [f"-DPB_F={f}"]) for f in (1, 2, 3, 4)] + \
        [("builder", "builder.cpp", []), ("common", "common.cpp", []), ("image_io", "image_io.cpp", [])]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def _deps():
    return [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(HERE, "..", "include", "probly_b200.h")]


def _compile_units(objdir: str, defs, force: bool, verbose: bool):
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for name, src, extra in UNITS:
        srcp = os.path.join(CSRC, src)
        if not os.path.exists(srcp):
            continue
        obj = os.path.join(objdir, name + ".o")
        if force or _newer(obj, [srcp] + _deps()):
            cmd = ["nvcc"] + NVCC_COMMON + [f"-D{d}" for d in defs] + extra + \
                  (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, srcp]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for rc, cmd in zip(ex.map(lambda c: subprocess.call(c), jobs), jobs):
                if rc != 0:
                    raise subprocess.CalledProcessError(rc, cmd)
    return [os.path.join(objdir, name + ".o") for name, src, _ in UNITS if os.path.exists(os.path.join(CSRC, src))], bool(jobs)


def build_variant(name: str, defs) -> str:
    """Tuning helper: the same library with extra -D flags, as _lib/libprobly_b200_<name>.so."""
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, f"libprobly_b200_{name}.so")
    objs, _ = _compile_units(os.path.join(OUT, "obj_" + name), list(defs), True, False)
    subprocess.check_call(["nvcc"] + LINK_FLAGS + ["-o", out] + objs)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    objs, changed = _compile_units(os.path.join(OUT, "obj"), [], force, verbose)
    if changed or not os.path.exists(LIB):
        subprocess.check_call(["nvcc"] + LINK_FLAGS + ["-o", LIB] + objs)
    wsrc = os.path.join(CSRC, "workload.cpp")
    if os.path.exists(wsrc) and (force or _newer(WLIB, [wsrc])):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-o", WLIB, wsrc])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
