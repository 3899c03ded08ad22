"""CPU tier: the oracle (oracle/oracle.cpp) against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py), and -- when oracle/_ref is present -- against the reference itself."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, random_soup
from oracle.bindings import OracleScene, fnv1a64_pairs, ref_available


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_tree_matches_reference_golden(name, meshes, golden, oracle_lib):
    P, I, N = meshes[name]
    g = golden["meshes"][name]
    s = OracleScene(oracle_lib); s.add_mesh(P, I, N); s.build()
    st = s.stats()
    for k in ("pairs", "nodes", "leaves", "max_depth", "min_leaf_depth", "overflow_leaves", "max_leaf_faces"):
        assert st[k] == g[k], k
    code, mesh, tri = s.pairs()
    assert zlib.crc32(code.tobytes()) == g["crc32_codes"] and zlib.crc32(tri.tobytes()) == g["crc32_tris"]
    assert "%016x" % fnv1a64_pairs(code, tri) == g["fnv1a64_pairs"]
    # radix-tree topology: a binary tree over L distinct codes has exactly L-1 two-child nodes
    assert st["branches"] == st["leaves"] - 1
    s.close()


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_trace_matches_reference_golden(name, meshes, golden, oracle_lib):
    P, I, N = meshes[name]
    s = OracleScene(oracle_lib); s.add_mesh(P, I, N); s.build()
    for res, g in golden["meshes"][name]["trace"].items():
        W, H = [int(x) for x in res.split("x")]
        rays = oracle_lib.initial_rays(W, H)
        r = s.trace_camera(rays, [0, 0, -2.1], np.eye(3, dtype=np.float32), 0, nthreads=os.cpu_count() or 1)
        assert int((r["tri"] != 0xFFFFFFFF).sum()) == g["hits"]
        assert zlib.crc32(r["pixels"].tobytes()) == g["crc32_pixels"]
        if res == "200x200":
            px = np.load(os.path.join(GOLDEN, "pixels_%s_200x200.npz" % name))["pixels"]
            assert np.array_equal(px.reshape(-1), r["pixels"])
    s.close()


def test_two_mesh_scene_matches_golden(meshes, golden, oracle_lib):
    P, I, N = meshes["f16"]
    a, b = golden["f16_material_groups"]
    s = OracleScene(oracle_lib); s.add_mesh(P, I[:a], N); s.add_mesh(P, I[a:], N); s.build()
    code, mesh, tri = s.pairs()
    g = golden["f16_two_meshes"]
    assert code.size == g["pairs"] and zlib.crc32(code.tobytes()) == g["crc32_codes"]
    assert zlib.crc32(mesh.tobytes()) == g["crc32_mesh"] and zlib.crc32(tri.tobytes()) == g["crc32_tris"]
    r = s.trace_camera(oracle_lib.initial_rays(200, 200), [0, 0, -2.1], np.eye(3, dtype=np.float32))
    assert zlib.crc32(r["pixels"].tobytes()) == g["crc32_pixels_200x200"]


def test_primitive_known_answers(golden, oracle_lib):
    L = oracle_lib.lib
    f = lambda *a: np.array(a, np.float32)
    dd = f(-1, -1, -1); dot = np.float32(dd[0] * dd[0] + dd[1] * dd[1]) + np.float32(dd[2] * dd[2])
    inv = (np.float32(1) / (dd * (np.float32(1) / np.sqrt(dot))).astype(np.float32)).astype(np.float32)
    k = golden["kat"]
    assert int(np.float32(L.orc_box_ray_nozero(f(0, 0, -3), f(1, 1, -2), f(0, 0, -2.3), inv)).view(np.uint32)) == k["box_ray_nozero_bits"] == 0x80000000
    assert int(np.float32(L.orc_box_ray(f(0, 0, -3), f(1, 1, -2), f(0, 0, -2.3), inv)).view(np.uint32)) == k["box_ray_bits"] == 0
    u, v = C.c_float(), C.c_float()
    t = L.orc_tri_intersect(f(0, 0, -2.1), f(0, 0, 1), f(-1, -1, 1.56), f(0, 1, 1.56), f(1, -1, 1.56), C.byref(u), C.byref(v))
    assert int(np.float32(t).view(np.uint32)) == k["tri_intersect"]["t_bits"] and u.value == 0.5 and v.value == 0.25
    assert abs(t - 3.65999985) < 1e-6


def test_primitive_vectors(oracle_lib):
    L = oracle_lib.lib
    z = np.load(os.path.join(GOLDEN, "primitives.npz"))
    n = z["cen"].shape[0]
    u, v = C.c_float(), C.c_float()
    for i in range(n):
        assert L.orc_tri_box_overlap(z["cen"][i].copy(), z["half"][i].copy(), np.ascontiguousarray(z["tri"][i].reshape(-1))) == z["sat"][i]
        t = L.orc_tri_intersect(z["org"][i].copy(), z["dir"][i].copy(), z["tri"][i, 0].copy(), z["tri"][i, 1].copy(), z["tri"][i, 2].copy(), C.byref(u), C.byref(v))
        assert np.float32(t).view(np.uint32) == z["t"][i].view(np.uint32)
        if t != np.finfo(np.float32).max:
            assert u.value == z["u"][i] and v.value == z["v"][i]
        with np.errstate(divide="ignore"):
            iv = (np.float32(1) / z["dir"][i]).astype(np.float32)
        b = L.orc_box_ray((z["cen"][i] - z["half"][i]).astype(np.float32), (z["cen"][i] + z["half"][i]).astype(np.float32), z["org"][i].copy(), iv)
        assert np.float32(b).view(np.uint32) == z["box"][i].view(np.uint32)
    assert 0.05 < z["sat"].mean() < 0.95 and (z["t"] != np.finfo(np.float32).max).sum() > 20


def test_hash_keys(golden, oracle_lib):
    L = oracle_lib.lib
    h = golden["hash"]
    for k, want in h["bmHash"]:
        assert L.orc_hash(k) == want
    for x, y, z, want in h["bmHash3"]:
        assert L.orc_hash3(x, y, z) == want
    for x, want in h["bmMap"]:
        assert L.orc_map(x) == want
    # survey KATs
    for k, want in {0: 0, 1: 1025, 2: 2050, 255: 0, 256: 769, 65535: 0, 65536: 513, 0xFFFFFFFF: 0, 4294966296: 22293, 12345678: 54892}.items():
        assert L.orc_hash(k) == want
    assert L.orc_hash3(1, 2, 3) == 6150
    z = np.load(os.path.join(GOLDEN, "hashkeys.npz"))
    for p, hk, c in zip(z["pts"], z["hash"], z["cells"]):
        cc = [L.orc_map(float(x)) for x in p]
        assert cc == list(c)
        assert L.orc_hash3(*[int(np.uint32(np.int32(x))) for x in cc]) == hk


def test_initial_rays(golden, oracle_lib):
    assert zlib.crc32(oracle_lib.initial_rays(200, 200).tobytes()) == golden["initial_rays_200x200_crc32"]
    assert zlib.crc32(oracle_lib.initial_rays(1280, 720).tobytes()) == golden["initial_rays_1280x720_crc32"]


def test_edge_cases(oracle_lib):
    # empty scene: every ray misses
    s = OracleScene(oracle_lib); s.build()
    assert s.stats()["pairs"] == 0
    r = s.trace_camera(oracle_lib.initial_rays(8, 8), [0, 0, -2.1], np.eye(3, dtype=np.float32))
    assert (r["tri"] == 0xFFFFFFFF).all() and (r["pixels"] == 255 << 8).all()
    # triangle outside the world box is dropped; degenerate triangle is kept where its point lies
    s = OracleScene(oracle_lib)
    P = np.array([[40, 40, 40], [41, 40, 40], [40, 41, 40], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]], np.float32)
    s.add_mesh(P, np.array([0, 1, 2, 3, 4, 5], np.uint32)); s.build()
    code, mesh, tri = s.pairs()
    assert (tri == 1).all() and code.size >= 1
    # leaf face cap (BuildTree.cuh:17): 300 copies of one small triangle -> every leaf keeps the first 256
    tri1 = np.array([[0.001, 0.001, 0.001], [0.01, 0.001, 0.001], [0.001, 0.01, 0.001]], np.float32)
    s = OracleScene(oracle_lib); s.add_mesh(np.tile(tri1, (300, 1)), np.arange(900, dtype=np.uint32)); s.build()
    st = s.stats()
    assert st["overflow_leaves"] == st["leaves"] >= 1 and st["max_leaf_faces"] == 300
    code, mesh, tri = s.pairs()
    assert tri.max() == 255


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (no /root/reference on this machine)")
def test_oracle_equals_reference_on_random_soups(oracle_lib):
    from oracle.bindings import RefLib, RefScene
    R = RefLib()
    rng = np.random.default_rng(3)
    for n, ext, edge in ((3000, 3.0, 0.05), (40, 25.0, 0.4)):   # stay inside the reference capacity of 65,536 leaves
        P, I = random_soup(rng, n, ext, edge)
        N = np.tile(np.array([[0, 0, 1]], np.float32), (P.shape[0], 1))
        a = RefScene(R); a.add_mesh(P, I, N); a.build()
        b = OracleScene(oracle_lib); b.add_mesh(P, I, N); b.build()
        ca, ma, ta = a.pairs(); cb, mb, tb = b.pairs()
        assert np.array_equal(ca, cb) and np.array_equal(ta, tb) and a.stats()["nodes"] == b.stats()["nodes"]
        eye = [0.1, 0.2, -ext - 1.0]
        err, px, _ = a.trace(96, 96, eye, np.eye(3, dtype=np.float32))
        r = b.trace_camera(oracle_lib.initial_rays(96, 96), eye, np.eye(3, dtype=np.float32))
        assert err == 0 and np.array_equal(px.reshape(-1), r["pixels"])
        a.close(); b.close()


def test_threaded_build_equals_the_serial_insertion(oracle_lib, meshes):
    """orc_scene_build_mt (used for the multi-million-triangle parity tests and the bench's CPU baseline) against the serial insertion that
    restates the reference: identical pair lists, two-child nodes, statistics -- leaf cap included -- and traced hits in both modes."""
    from raytracercuda_b200 import synth
    rng = np.random.default_rng(5)
    c = rng.uniform(-0.05, 0.05, (1500, 1, 3)).astype(np.float32)
    dense = np.ascontiguousarray((c + rng.uniform(-0.02, 0.02, (1500, 3, 3)).astype(np.float32)).reshape(-1, 3))       # leaves over the 256-face cap
    cases = [("f16",) + meshes["f16"], ("suzanne",) + meshes["suzanne"], ("soup",) + synth.soup(30000, extent=3.0, edge=0.05) + (None,),
             ("dense", dense, np.arange(4500, dtype=np.uint32).reshape(-1, 3), None)]
    rays = synth.incoherent_rays(4000, extent=3.0)
    for name, P, I, N in cases:
        a = OracleScene(oracle_lib); a.add_mesh(P, I, N); a.build()
        b = OracleScene(oracle_lib); b.add_mesh(P, I, N); b.add_mesh(P[::-1].copy(), I, N)          # two meshes: the (mesh, triangle) order matters
        a2 = OracleScene(oracle_lib); a2.add_mesh(P, I, N); a2.add_mesh(P[::-1].copy(), I, N); a2.build()
        one = OracleScene(oracle_lib); one.add_mesh(P, I, N); one.build(5)
        b.build(3)
        for x, y in ((a, one), (a2, b)):
            assert x.stats() == y.stats(), name
            assert all(np.array_equal(p, q) for p, q in zip(x.pairs(), y.pairs())), name
            assert all(np.array_equal(p, q) for p, q in zip(x.branches(), y.branches())), name
            for mode in (0, 1):
                rx = x.trace_rays(rays[:, :3], rays[:, 3:], mode, 2); ry = y.trace_rays(rays[:, :3], rays[:, 3:], mode, 2)
                assert all(np.array_equal(rx[k], ry[k]) for k in ("t", "u", "v", "mesh", "tri", "tie", "pixels")), (name, mode)
        if name == "dense":
            assert a.stats()["overflow_leaves"] > 0
        for s_ in (a, b, a2, one):
            s_.close()
