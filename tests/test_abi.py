"""CPU tier: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol the headers declare, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    names = []
    for hdr in ("beam_b200.h", "beam_host_c.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b((?:bm|beam)[A-Z]\w*)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(beam_lib):
    from raytracercuda_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 50
    missing = [n for n in declared if not hasattr(beam_lib, n)]
    assert not missing, missing
    assert sorted(_lib.ABI_SYMBOLS) == declared, "raytracercuda_b200/_lib.py ABI_SYMBOLS out of sync with include/*.h"


def test_sm100a_code_is_in_the_library(beam_lib):
    import subprocess
    from raytracercuda_b200._lib import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_reference_struct_sizes(beam_lib):
    from raytracercuda_b200 import _lib
    assert C.sizeof(_lib.bmStaticMeshData) == 136 and C.sizeof(_lib.bmVec3) == 12 and C.sizeof(_lib.bmMat3) == 36
    assert (beam_lib.bmGetFaceSize(), beam_lib.bmGetNodeSize(), beam_lib.bmGetFacePtrSize(), beam_lib.bmGetMaterialSize()) == (32, 32, 8, 4)
    # SURVEY.md section 8c struct sizes (LP64) behind the HASH and PROGRESSIVE seams
    L = beam_lib
    assert (L.bmGetCellSize(), L.bmGetMaxHashElements()) == (8208, 65536)
    assert (L.bmGetStoreSize(), L.bmGetFaceClusterSize(), L.bmGetBvhNodeSize(), L.bmGetRaySize(), L.bmGetRayBoxSize(), L.bmGetRayFaceSize(),
            L.bmRayFaceHitClusterSize()) == (16, 512, 32, 24, 16, 16, 392)


def test_no_cpu_fallback(beam_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device failure mode is exercised on the CPU tier")
    h = C.c_void_p()
    rc = beam_lib.bmTreeCreate(C.byref(h), None, None, 0.0, 0)
    assert rc == 100 and b"no CPU path" in beam_lib.bmLastErrorString()
    assert not beam_lib.beamSceneCreate()
    from raytracercuda_b200 import BeamError
    from raytracercuda_b200.tree import Tree
    with pytest.raises((BeamError, RuntimeError, AssertionError)):
        Tree("cuda:0")


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under raytracercuda_b200/ may import, load or name it"""
    pkg = os.path.join(ROOT, "raytracercuda_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle.bindings" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f


def test_reference_arm_contract_runs_without_a_gpu():
    """`bench.py --impl reference` is the CPU arm: it must run on a box without CUDA and print exactly ONE JSON line carrying the contract's keys"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--reference-rays", "4000", "--workload", "c2small"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
