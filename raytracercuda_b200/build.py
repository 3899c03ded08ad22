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
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libbeam_b200.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
