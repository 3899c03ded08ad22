"""CPU tier: the product's host+device inline arithmetic (csrc/bm_math.cuh, bm_sat.cuh, bm_grid_host.h) compiled for the
CPU (tests/hostcheck) and checked against the oracle -- the same source the kernels execute."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_soup
from oracle.bindings import OracleScene

f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(ROOT, "tests", "hostcheck", "_hostcheck.so")
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src])
    H = C.CDLL(so)
    H.hc_pairs.argtypes = [f32p, u32p, C.c_uint32, f32p, f32p, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    H.hc_pairs.restype = C.c_uint64
    H.hc_pairs_flat.argtypes = [f32p, u32p, C.c_uint32, f32p, f32p, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    H.hc_pairs_flat.restype = C.c_uint64
    H.hc_grid.argtypes = [f32p, f32p, C.c_float, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
    H.hc_encode.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]; H.hc_encode.restype = C.c_uint32
    H.hc_decode.argtypes = [C.c_int, C.c_uint32, u32p]
    H.hc_box_ray.argtypes = [f32p] * 4; H.hc_box_ray.restype = C.c_float
    H.hc_tri_intersect.argtypes = [f32p] * 5 + [C.POINTER(C.c_float)] * 2; H.hc_tri_intersect.restype = C.c_float
    H.hc_tri_pretest.argtypes = [f32p, f32p, f32p, C.c_uint64, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")]; H.hc_tri_pretest.restype = None
    H.hc_wide_trace.argtypes = [f32p, u32p, C.c_uint32, f32p, f32p, C.c_float, C.c_uint32, C.c_uint32, f32p, C.c_uint64, C.c_int, u32p, f32p,
                                np.ctypeslib.ndpointer(np.uint64, flags="C")]
    H.hc_hash.argtypes = [C.c_uint32]; H.hc_hash.restype = C.c_uint32
    H.hc_map.argtypes = [C.c_float]; H.hc_map.restype = C.c_int32
    return H


def _pairs(H, P, I, wmin=(-30, -30, -30), wmax=(30, 30, 30)):
    I = np.ascontiguousarray(I, np.uint32).reshape(-1)
    a = np.array(wmin, np.float32); b = np.array(wmax, np.float32); D = C.c_int(0)
    n = H.hc_pairs(P, I, I.size // 3, a, b, 0.03, 38, None, None, 0, C.byref(D))
    assert n != 2**64 - 1
    code = np.zeros(n, np.uint32); tri = np.zeros(n, np.uint32)
    H.hc_pairs(P, I, I.size // 3, a, b, 0.03, 38, code.ctypes.data, tri.ctypes.data, n, C.byref(D))
    return code, tri, D.value


def _pairs_flat(H, P, I, wmin=(-30, -30, -30), wmax=(30, 30, 30)):
    I = np.ascontiguousarray(I, np.uint32).reshape(-1)
    a = np.array(wmin, np.float32); b = np.array(wmax, np.float32); cand = C.c_uint64(0)
    n = H.hc_pairs_flat(P, I, I.size // 3, a, b, 0.03, 38, None, None, 0, C.byref(cand))
    code = np.zeros(n, np.uint32); tri = np.zeros(n, np.uint32)
    H.hc_pairs_flat(P, I, I.size // 3, a, b, 0.03, 38, code.ctypes.data, tri.ctypes.data, n, C.byref(cand))
    return code, tri, cand.value


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_flat_keygen_equals_oracle_on_meshes(name, hc, meshes, oracle_lib):
    P, I, N = meshes[name]
    s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
    oc, om, ot = s.pairs()
    code, tri, cand = _pairs_flat(hc, P, I)
    assert cand > I.shape[0]
    assert np.array_equal(code, oc) and np.array_equal(tri, ot)


def test_flat_keygen_equals_oracle_near_split_planes(hc, oracle_lib):
    """triangles hugging the large ancestors' split planes (x=0, y=0, z=0, +-15, +-7.5 ...), where the float SAT of the big boxes
    is least trustworthy, plus thin slivers and axis-aligned triangles"""
    rng = np.random.default_rng(21)
    planes = np.array([0.0, 15.0, -15.0, 7.5, -7.5, 3.75, 22.5, -0.029296875, 0.05859375], np.float32)
    parts = []
    for k in range(6000):
        c = rng.uniform(-29, 29, 3)
        ax = rng.integers(0, 3)
        c[ax] = planes[rng.integers(0, planes.size)] + rng.choice([0.0, 1e-7, -1e-7, 1e-5, -1e-5, 1e-3, -1e-3, 4e-3])
        e = rng.uniform(-0.02, 0.02, (3, 3))
        kind = k % 4
        if kind == 1: e[:, ax] *= 1e-3                 # nearly parallel to the plane
        if kind == 2: e[:, ax] = 0.0                   # exactly in a plane parallel to it
        if kind == 3: e[:, (ax + 1) % 3] = 0.0         # edge vectors with zero components: degenerate cross axes
        parts.append((c + e).astype(np.float32))
    P = np.concatenate(parts, 0); I = np.arange(P.shape[0], dtype=np.uint32).reshape(-1, 3)
    s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
    oc, om, ot = s.pairs()
    code, tri, cand = _pairs_flat(hc, P, I)
    assert np.array_equal(code, oc) and np.array_equal(tri, ot)
    dcode, dtri, D = _pairs(hc, P, I)
    assert np.array_equal(dcode, oc) and np.array_equal(dtri, ot)


def test_flat_keygen_equals_oracle_on_soups(hc, oracle_lib):
    rng = np.random.default_rng(22)
    for n, ext, edge in ((8000, 29.0, 0.05), (3000, 2.0, 0.004), (1500, 10.0, 0.08), (200, 31.0, 1.5)):
        P, I = random_soup(rng, n, ext, edge)
        s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
        oc, om, ot = s.pairs()
        code, tri, cand = _pairs_flat(hc, P, I)
        assert np.array_equal(code, oc) and np.array_equal(tri, ot), (n, ext, edge)


@pytest.mark.parametrize("name", ["f16", "suzanne"])
def test_descent_equals_oracle_on_meshes(name, hc, meshes, oracle_lib):
    P, I, N = meshes[name]
    s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
    oc, om, ot = s.pairs()
    code, tri, D = _pairs(hc, P, I)
    assert D == 31 and np.array_equal(code, oc) and np.array_equal(tri, ot)


def test_descent_equals_oracle_on_soups(hc, oracle_lib):
    rng = np.random.default_rng(1)
    for n, ext, edge in ((8000, 29.0, 0.05), (200, 31.0, 1.5)):
        P, I = random_soup(rng, n, ext, edge)
        s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
        oc, om, ot = s.pairs()
        code, tri, D = _pairs(hc, P, I)
        assert np.array_equal(code, oc) and np.array_equal(tri, ot)


def test_grid_default_world(hc):
    k = (C.c_int * 3)(); D = C.c_int()
    tx = np.zeros(2049, np.float32); ty = np.zeros(1025, np.float32); tz = np.zeros(1025, np.float32)
    a = np.full(3, -30, np.float32); b = np.full(3, 30, np.float32)
    assert hc.hc_grid(a, b, 0.03, 38, k, C.byref(D), tx.ctypes.data, ty.ctypes.data, tz.ctypes.data) == 1
    assert list(k) == [11, 10, 10] and D.value == 31               # every leaf at depth 31, 2048x1024x1024 cells
    assert np.array_equal(tx, (-30 + np.arange(2049) * (60.0 / 2048)).astype(np.float32))      # boundaries are exact floats
    assert np.array_equal(ty, (-30 + np.arange(1025) * (60.0 / 1024)).astype(np.float32))
    assert hc.hc_grid(np.full(3, -1000, np.float32), np.full(3, 1000, np.float32), 0.03, 38, k, C.byref(D), None, None, None) == 0


def test_path_code_roundtrip(hc):
    rng = np.random.default_rng(2)
    out = np.zeros(3, np.uint32)
    for depth in range(1, 33):
        nb = [(depth + 2 - a) // 3 for a in range(3)]
        for _ in range(50):
            xyz = [int(rng.integers(0, 1 << nb[a])) if nb[a] else 0 for a in range(3)]
            p = hc.hc_encode(depth, *xyz)
            assert p < (1 << depth) if depth < 32 else True
            hc.hc_decode(depth, p, out)
            assert list(out) == xyz
            # bit d of the path (MSB first) is bit of axis d%3, most significant first
            bits = []
            cnt = [0, 0, 0]
            for d in range(depth):
                a = d % 3
                bits.append((xyz[a] >> (nb[a] - 1 - cnt[a])) & 1); cnt[a] += 1
            assert p == int("".join(map(str, bits)), 2)


def test_primitives_equal_golden(hc):
    z = np.load(os.path.join(GOLDEN, "primitives.npz"))
    u, v = C.c_float(), C.c_float()
    for i in range(z["cen"].shape[0]):
        t = hc.hc_tri_intersect(z["org"][i].copy(), z["dir"][i].copy(), z["tri"][i, 0].copy(), z["tri"][i, 1].copy(), z["tri"][i, 2].copy(), C.byref(u), C.byref(v))
        assert np.float32(t).view(np.uint32) == z["t"][i].view(np.uint32)
        with np.errstate(divide="ignore"):
            iv = (np.float32(1) / z["dir"][i]).astype(np.float32)
        b = hc.hc_box_ray((z["cen"][i] - z["half"][i]).astype(np.float32), (z["cen"][i] + z["half"][i]).astype(np.float32), z["org"][i].copy(), iv)
        assert np.float32(b).view(np.uint32) == z["box"][i].view(np.uint32)


def test_hash_keys(hc, golden):
    for k, want in golden["hash"]["bmHash"]:
        assert hc.hc_hash(k) == want
    for x, want in golden["hash"]["bmMap"]:
        assert hc.hc_map(x) == want


def _wide_vs_oracle(hc, oracle_lib, P, I, N, rays, what, max_faces=256, wmin=(-30, -30, -30), wmax=(30, 30, 30), exact=True, modes=(0, 1)):
    """complete traversals through the product's 64-byte wide records (encoder + traversal step of csrc/bm_wide.cuh, compiled for the CPU)
    against the oracle: same winner and bit-equal t in both trace modes, and every decoded slot box encloses the exact one"""
    P = np.ascontiguousarray(P, np.float32); Iflat = np.ascontiguousarray(I, np.uint32).reshape(-1); rays = np.ascontiguousarray(rays, np.float32)
    s = OracleScene(oracle_lib, wmin=wmin, wmax=wmax, max_faces=max_faces); s.add_mesh(P, Iflat.reshape(-1, 3), N); s.build()
    out = {}
    for mode in modes:
        tri = np.zeros(rays.shape[0], np.uint32); t = np.zeros(rays.shape[0], np.float32); st = np.zeros(4, np.uint64)
        assert hc.hc_wide_trace(P, Iflat, Iflat.size // 3, np.array(wmin, np.float32), np.array(wmax, np.float32), 0.03, 38, max_faces, rays, rays.shape[0], mode, tri, t, st) == 0
        r = s.trace_rays(rays[:, :3], rays[:, 3:], mode, 4)
        ok = r["tie"] == 0
        agree = (tri[ok] == r["tri"][ok]).mean()
        assert agree == 1.0 if exact else agree >= 0.9999, (what, mode, agree)
        both = ok & (tri == r["tri"]) & (tri != 0xFFFFFFFF)
        assert np.array_equal(t[both], r["t"][both]), (what, mode)
        assert st[2] == 0, "decoded culling boxes must enclose the exact ones"
        out[mode] = (st[0] / rays.shape[0], r["node_visits"] / rays.shape[0], int(st[3]))
    s.close()
    return out


def _rays_from(eye, dirs):
    return np.concatenate([np.tile(np.asarray(eye, np.float32), (dirs.shape[0], 1)), dirs], 1).astype(np.float32)


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_wide_nodes_trace_like_the_oracle_on_meshes(name, hc, meshes, oracle_lib):
    from raytracercuda_b200 import synth
    P, I, N = meshes[name]
    v = _wide_vs_oracle(hc, oracle_lib, P, I, N, _rays_from([0, 0, -2.1], oracle_lib.initial_rays(96, 96)), name + " camera")
    assert v[0][0] < 0.25 * v[0][1]                    # a fraction of the reference's node visits
    _wide_vs_oracle(hc, oracle_lib, P, I, N, synth.incoherent_rays(6000, extent=1.5), name + " incoherent")
    assert v[0][2] == 0 and v[1][2] == 0               # no axis had to fall back to "unbounded"


def test_wide_nodes_trace_like_the_oracle_on_soups(hc, oracle_lib):
    from raytracercuda_b200 import synth
    P, I = synth.soup(20000, extent=3.0, edge=0.05)
    _wide_vs_oracle(hc, oracle_lib, P, I, None, synth.incoherent_rays(12000, extent=3.2), "soup")
    # axis-parallel rays, in-plane rays, origins exactly on cell boundaries (the NaN / inf corners of the reference's slab test)
    rng = np.random.default_rng(4)
    n = 4000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    o[::3, 0] = np.round(o[::3, 0] / 0.029296875) * 0.029296875; o[1::5, 1] = 0
    d = np.zeros((n, 3), np.float32); d[np.arange(n), rng.integers(0, 3, n)] = rng.choice([-1, 1], n)
    m = d[::7].shape[0]
    d[::7] = rng.normal(size=(m, 3)); d[::7, 2] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
    _wide_vs_oracle(hc, oracle_lib, P, I, None, np.concatenate([o, d], 1), "soup, degenerate rays")
    # triangles far larger than the cells (quantisation grids anchored at the world box)
    P, I = synth.soup(600, extent=22.0, edge=3.5)
    _wide_vs_oracle(hc, oracle_lib, P, I, None, synth.incoherent_rays(4000, extent=29), "big triangles")
    # ... some sticking out of the world box: the reference's semantics only know the part inside (a closest hit outside the world has no
    # leaf cell to be found in), so only the reference mode is compared exactly there
    P, I = synth.soup(600, extent=28.0, edge=3.5)
    _wide_vs_oracle(hc, oracle_lib, P, I, None, synth.incoherent_rays(4000, extent=29), "big triangles across the world box", modes=(0,))
    # leaves above the reference's face cap, a non-default world, and the table (non-closed-form) grid
    c = rng.uniform(-0.05, 0.05, (1200, 1, 3)).astype(np.float32)
    dense = np.ascontiguousarray((c + rng.uniform(-0.02, 0.02, (1200, 3, 3)).astype(np.float32)).reshape(-1, 3))
    _wide_vs_oracle(hc, oracle_lib, dense, np.arange(3600, dtype=np.uint32), None, synth.incoherent_rays(4000, extent=0.3), "dense, capped leaves", max_faces=64, modes=(0,))      # (closest mode is defined without the cap)
    P, I = synth.soup(5000, extent=2.5, edge=0.05)
    _wide_vs_oracle(hc, oracle_lib, P + np.float32(1.0), I, None, synth.incoherent_rays(5000, extent=2.5) + np.array([1, 1, 1, 0, 0, 0], np.float32), "other world",
                    wmin=(-3.1, -2.7, -2.2), wmax=(4.9, 5.3, 5.8))
    os.environ["HC_NO_UNIFORM"] = "1"
    try:
        _wide_vs_oracle(hc, oracle_lib, P, I, None, synth.incoherent_rays(3000, extent=2.5), "table grid")
    finally:
        del os.environ["HC_NO_UNIFORM"]


def test_triangle_pretest_never_rejects_what_the_exact_test_accepts(hc):
    """tri_reject (bm_math.cuh): the fused, division-free pre-test the trace kernels run in front of the reference-order triangle test.
    It must never answer "no hit" where tri_intersect accepts; checked on rays aimed AT edges, vertices and interiors of random triangles
    (targets computed in double, so the exact test's u, v land within rounding of 0, 1 and u + v = 1), on near-parallel rays, on huge,
    tiny and needle triangles, and on far origins -- and it has to be worth running: it must catch most of the misses."""
    rng = np.random.default_rng(2024)
    total = np.zeros(4, np.uint64)
    for scale, far in ((1.0, 1.0), (1e-3, 1.0), (40.0, 1.0), (1.0, 50.0), (0.02, 30.0)):
        n = 400_000
        tri = (rng.standard_normal((n, 3, 3)) * scale).astype(np.float64)
        kind = rng.integers(0, 6, n)
        needle = kind == 5
        tri[needle, 2] = tri[needle, 1] + (tri[needle, 2] - tri[needle, 1]) * 1e-4           # nearly degenerate
        # target point on the triangle's plane: barycentric (u, v) chosen on / near the acceptance boundary
        u = rng.random(n); v = rng.random(n) * (1 - u)
        edge = rng.integers(0, 4, n)
        on = kind <= 2
        u = np.where(on & (edge == 0), 0.0, u); v = np.where(on & (edge == 1), 0.0, v)
        v = np.where(on & (edge == 2), 1.0 - u, v)
        u = np.where(on & (edge == 3), np.round(u), u); v = np.where(on & (edge == 3), np.where(u == 1, 0.0, np.round(v)), v)     # vertices
        miss = kind == 3
        u = np.where(miss, rng.standard_normal(n) * 2, u); v = np.where(miss, rng.standard_normal(n) * 2, v)
        target = tri[:, 0] + u[:, None] * (tri[:, 1] - tri[:, 0]) + v[:, None] * (tri[:, 2] - tri[:, 0])
        org = target + rng.standard_normal((n, 3)) * scale * far
        par = kind == 4                                                                                  # rays (nearly) inside the plane
        inplane = (tri[:, 1] - tri[:, 0]) * rng.standard_normal((n, 1)) + (tri[:, 2] - tri[:, 0]) * rng.standard_normal((n, 1))
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        org = np.where(par[:, None], target + inplane + nrm * (rng.standard_normal((n, 1)) * 1e-7), org)
        d = target - org
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)
        counts = np.zeros(4, np.uint64)
        hc.hc_tri_pretest(np.ascontiguousarray(org, np.float32).reshape(-1), np.ascontiguousarray(d, np.float32).reshape(-1),
                          np.ascontiguousarray(tri, np.float32).reshape(-1), n, counts)
        assert counts[3] == 0, "pre-test rejected %d accepted hits at scale %g / distance %g" % (counts[3], scale, far)
        assert counts[0] > 0.2 * n and counts[1] > 0.1 * n                 # both outcomes are exercised
        total += counts
    # ordinary misses (what a leaf's face loop mostly sees: the ray passes somewhere near the triangle): nearly all must be caught
    n = 400_000
    tri = rng.standard_normal((n, 3, 3)).astype(np.float32)
    org = (rng.standard_normal((n, 3)) * 3).astype(np.float32)
    d = (tri.mean(axis=1) + rng.standard_normal((n, 3)) * 1.5 - org).astype(np.float64)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    counts = np.zeros(4, np.uint64)
    hc.hc_tri_pretest(np.ascontiguousarray(org).reshape(-1), np.ascontiguousarray(d).reshape(-1), np.ascontiguousarray(tri).reshape(-1), n, counts)
    assert counts[3] == 0 and counts[0] > 1000
    assert counts[2] >= 0.999 * counts[1], "the pre-test catches only %.2f %% of ordinary misses" % (100.0 * counts[2] / counts[1])
