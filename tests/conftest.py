import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _build_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return so


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.fixture(scope="session")
def meshes():
    z = np.load(os.path.join(GOLDEN, "meshes.npz"))
    return {n: (z[n + "_pos"], z[n + "_idx"], z[n + "_nrm"]) for n in ("f16", "suzanne", "bunny")}


@pytest.fixture(scope="session")
def oracle_lib():
    _build_oracle()
    from oracle.bindings import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def beam_lib():
    """the product library, (re)built in-tree if stale"""
    from raytracercuda_b200 import build, _lib
    build.build_library()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test ran without a CUDA device; the product has no CPU path")
    return "cuda:0"


def random_soup(rng, n, extent=3.0, edge=0.05):
    c = rng.uniform(-extent, extent, (n, 1, 3)).astype(np.float32)
    P = (c + rng.uniform(-edge, edge, (n, 3, 3)).astype(np.float32)).reshape(-1, 3)
    I = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
    return np.ascontiguousarray(P), I
