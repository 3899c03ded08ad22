"""Builds raytracercuda_b200/libbeam_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

`python -m raytracercuda_b200.build` or `__graft_entry__.build()`.  A CMakeLists.txt with the same sources and flags sits at
the repository root for CMake users; this script is what the tests and the driver use (no cmake dependency at run time).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbeam_b200.so")
SOURCES = ["scan.cu", "sort.cu", "keygen.cu", "hierarchy.cu", "trace.cu", "bricks.cu", "rays.cu", "tree.cu", "host.cu", "objload.cpp"]
# -fmad=false: keys, traversal decisions and t must round like the reference's unfused host arithmetic (DESIGN.md "Rounding")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-ccbin", "g++", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "beam_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """every source is compiled to its own object (in parallel; an object is rebuilt when its source, any header of csrc/ or include/, or
    this script is newer), then linked: a one-file change rebuilds in seconds"""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = os.path.join(HERE, "_obj")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    inc = os.path.join(HERE, "..", "include")
    headers += [os.path.join(inc, f) for f in os.listdir(inc)] + [__file__]
    newest_header = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(name):
        src = os.path.join(CSRC, name); obj = os.path.join(obj_dir, name + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj, 0, ""
        r = subprocess.run(["nvcc"] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]))
    for _, rc, out in results:
        if verbose or rc != 0:
            sys.stderr.write(out)
    if any(rc != 0 for _, rc, _ in results):
        raise RuntimeError("nvcc failed building libbeam_b200.so")
    r = subprocess.run(["nvcc"] + NVCC_FLAGS + ["-o", LIB] + [o for o, _, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libbeam_b200.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
