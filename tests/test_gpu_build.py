"""GPU tier: the CUDA build path (keys -> sort -> leaves -> bottom-up radix tree with tight boxes) through the C ABI, against the oracle.
Bar (BASELINE.json north_star): keys, sorted primitive order and topology bit-exact; node bounds within 1e-5 relative."""
import os
import zlib

import numpy as np
import pytest

from conftest import random_soup
from oracle.bindings import OracleScene, fnv1a64_pairs

pytestmark = pytest.mark.gpu


def _gpu_tree(dev, parts, **kw):
    from raytracercuda_b200.tree import Tree
    t = Tree(dev, **kw)
    for P, I in parts:
        t.add_mesh(P, I)
    t.build()
    return t


def _oracle(oracle_lib, parts, **kw):
    s = OracleScene(oracle_lib, **kw)
    for P, I in parts:
        s.add_mesh(P, I)
    s.build()
    return s


def _check_tree_against_oracle(t, s, parts):
    st, so = t.stats(), s.stats()
    assert st["leaves"] == so["leaves"]
    assert st["max_leaf_faces"] == so["max_leaf_faces"] and st["overflow_leaves"] == so["overflow_leaves"]
    code, tri = t.pairs()
    oc, om, ot = s.pairs()
    bases = t.mesh_bases()
    # the reference keeps at most 256 faces per leaf (BuildTree.cuh:17, BuildTree.cu:53-60) and so does the oracle's list;
    # the GPU list keeps every pair (the cap is applied by the reference-mode trace) -> cap it the same way to compare
    heads = np.flatnonzero(np.concatenate([[True], code[1:] != code[:-1]])) if code.size else np.zeros(0, np.int64)
    rank = np.arange(code.size) - np.repeat(heads, np.diff(np.concatenate([heads, [code.size]])))
    keep = rank < 256
    assert (~keep).sum() == 0 or so["overflow_leaves"] > 0
    assert np.array_equal(code[keep], oc), "leaf path codes differ"
    assert np.array_equal(tri[keep], bases[om] + ot), "sorted primitive order differs"
    # sortedness by (code, global triangle id): the property the stable sort must deliver
    key = (code.astype(np.uint64) << np.uint64(32)) | tri.astype(np.uint64)
    assert (key[1:] > key[:-1]).all()
    # topology: the two-child nodes of the reference tree == internal nodes of the radix tree, as (depth, prefix)
    plen, prefix, c0, c1 = t.topology()
    od, op = s.branches()
    mine = np.stack([plen, prefix], 1); mine = mine[np.lexsort((mine[:, 1], mine[:, 0]))]
    assert np.array_equal(mine, np.stack([od, op], 1)), "radix-tree topology differs from the reference's two-child nodes"
    # structural checks: every leaf and every internal node except the root is referenced exactly once; left child = smaller codes
    L, N = st["leaves"], st["nodes"]
    if N:
        refs = np.concatenate([c0, c1])
        leaf_refs = np.sort(refs[refs >= 0x80000000] & 0x7FFFFFFF); node_refs = np.sort(refs[refs < 0x80000000])
        root = t.root()
        assert np.array_equal(leaf_refs, np.arange(L)) and np.array_equal(node_refs, np.delete(np.arange(N), root))
        assert plen[root] == plen.min()
    else:
        assert t.root() == (0x80000000 if L == 1 else 0)
    return code, tri, plen, prefix, c0, c1


def _check_child_boxes(t, parts, code, tri, c0, c1, world_extent=60.0):
    """tight child boxes == union of the subtree's whole triangles (+ the documented pad of extent*2^-20)"""
    N = t.stats()["nodes"]
    if N == 0:
        return
    boxes = t.child_boxes()
    P = np.concatenate([p[np.asarray(i, np.int64).reshape(-1, 3)] for p, i in parts], 0)     # (T,3,3) in global id order
    tmin, tmax = P.min(1), P.max(1)
    heads = np.flatnonzero(np.concatenate([[True], code[1:] != code[:-1]]))
    L = heads.size
    lmin = np.minimum.reduceat(tmin[tri], heads, 0); lmax = np.maximum.reduceat(tmax[tri], heads, 0)
    nmin = np.full((N, 3), np.inf); nmax = np.full((N, 3), -np.inf)
    # children have larger indices than... not guaranteed; resolve by repeated relaxation from the leaves up
    def child_box(c):
        isleaf = c >= 0x80000000
        idx = (c & 0x7FFFFFFF).astype(np.int64)
        return np.where(isleaf[:, None], lmin[np.minimum(idx, L - 1)], nmin[np.minimum(idx, N - 1)]), np.where(isleaf[:, None], lmax[np.minimum(idx, L - 1)], nmax[np.minimum(idx, N - 1)])
    for _ in range(64):
        a0, b0 = child_box(c0); a1, b1 = child_box(c1)
        nm, nx = np.minimum(a0, a1), np.maximum(b0, b1)
        if np.array_equal(nm, nmin) and np.array_equal(nx, nmax):
            break
        nmin, nmax = nm, nx
    a0, b0 = child_box(c0); a1, b1 = child_box(c1)
    want = np.concatenate([a0, b0, a1, b1], 1)
    pad = world_extent * 2.0 ** -20
    tol = 1e-5 * world_extent          # "node bounds within 1e-5 relative" (relative to the world extent)
    assert np.isfinite(want).all()
    assert np.abs(boxes - want).max() <= pad * 1.01 + 1e-6 < tol
    # containment: the stored box is never tighter than the exact one
    assert (boxes[:, 0:3] <= want[:, 0:3]).all() and (boxes[:, 3:6] >= want[:, 3:6]).all()
    assert (boxes[:, 6:9] <= want[:, 6:9]).all() and (boxes[:, 9:12] >= want[:, 9:12]).all()


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_bundled_meshes_bit_exact(name, meshes, golden, oracle_lib, beam_lib, cuda_device):
    P, I, _ = meshes[name]
    parts = [(P, I)]
    t = _gpu_tree(cuda_device, parts); s = _oracle(oracle_lib, parts)
    code, tri, plen, prefix, c0, c1 = _check_tree_against_oracle(t, s, parts)
    g = golden["meshes"][name]          # and straight against the reference-generated fingerprints
    assert code.size == g["pairs"] and t.stats()["leaves"] == g["leaves"]
    assert zlib.crc32(code.tobytes()) == g["crc32_codes"] and zlib.crc32(tri.tobytes()) == g["crc32_tris"]
    assert "%016x" % fnv1a64_pairs(code, tri) == g["fnv1a64_pairs"]
    _check_child_boxes(t, parts, code, tri, c0, c1)
    t.close(); s.close()


def test_two_meshes_global_ids(meshes, golden, oracle_lib, beam_lib, cuda_device):
    P, I, _ = meshes["f16"]
    a, b = golden["f16_material_groups"]
    parts = [(P, I[:a]), (P, I[a:])]
    t = _gpu_tree(cuda_device, parts); s = _oracle(oracle_lib, parts)
    code, tri, *_ = _check_tree_against_oracle(t, s, parts)
    assert list(t.mesh_bases()) == [0, a] and t.stats()["meshes"] == 2
    assert zlib.crc32(code.tobytes()) == golden["f16_two_meshes"]["crc32_codes"]
    t.close(); s.close()


@pytest.mark.parametrize("n,ext,edge,seed", [(20000, 3.0, 0.05, 1), (50000, 29.0, 0.03, 2), (300, 31.0, 1.5, 3), (5000, 0.2, 0.2, 4)])
def test_random_soups_bit_exact(n, ext, edge, seed, oracle_lib, beam_lib, cuda_device):
    """small / spread-out / large-and-partly-outside-the-world / heavily overlapping (deep leaves) soups"""
    rng = np.random.default_rng(seed)
    P, I = random_soup(rng, n, ext, edge)
    parts = [(P, I)]
    t = _gpu_tree(cuda_device, parts); s = _oracle(oracle_lib, parts)
    code, tri, plen, prefix, c0, c1 = _check_tree_against_oracle(t, s, parts)
    _check_child_boxes(t, parts, code, tri, c0, c1)
    t.close(); s.close()


def test_edge_cases(oracle_lib, beam_lib, cuda_device):
    from raytracercuda_b200.tree import Tree
    # empty scene
    t = Tree(cuda_device); t.build()
    assert t.stats()["pairs"] == 0 and t.stats()["leaves"] == 0
    t.close()
    # everything outside the world box
    P = np.array([[40, 40, 40], [41, 40, 40], [40, 41, 40]], np.float32)
    t = _gpu_tree(cuda_device, [(P, np.array([[0, 1, 2]], np.uint32))])
    assert t.stats()["triangles"] == 1 and t.stats()["pairs"] == 0
    t.close()
    # a single tiny triangle -> a single leaf (root is a leaf), and a degenerate (point) triangle
    P = np.array([[0.501, 0.501, 0.501], [0.502, 0.501, 0.501], [0.501, 0.502, 0.501], [0.7, 0.7, 0.7], [0.7, 0.7, 0.7], [0.7, 0.7, 0.7]], np.float32)
    for I in (np.array([[0, 1, 2]], np.uint32), np.array([[0, 1, 2], [3, 4, 5]], np.uint32)):
        parts = [(P, I)]
        t = _gpu_tree(cuda_device, parts); s = _oracle(oracle_lib, parts)
        _check_tree_against_oracle(t, s, parts)
        t.close(); s.close()
    # leaf face cap: 300 copies of one triangle; the GPU keeps all pairs and reports the overflow like the reference counts it
    tri1 = np.array([[0.001, 0.001, 0.001], [0.01, 0.001, 0.001], [0.001, 0.01, 0.001]], np.float32)
    parts = [(np.tile(tri1, (300, 1)), np.arange(900, dtype=np.uint32).reshape(-1, 3))]
    t = _gpu_tree(cuda_device, parts); s = _oracle(oracle_lib, parts)
    st, so = t.stats(), s.stats()
    assert st["leaves"] == so["leaves"] == st["overflow_leaves"] == so["overflow_leaves"] and st["max_leaf_faces"] == 300
    code, tri = t.pairs()
    oc, om, ot = s.pairs()
    for c in np.unique(oc):
        assert np.array_equal(tri[code == c][:256], ot[oc == c])
    t.close(); s.close()
    # rebuild after reset reuses the handle
    rng = np.random.default_rng(9)
    t = Tree(cuda_device)
    for k in range(2):
        P, I = random_soup(rng, 2000, 2.0, 0.05)
        t.reset(); t.add_mesh(P, I); t.build()
        s = _oracle(oracle_lib, [(P, I)])
        _check_tree_against_oracle(t, s, [(P, I)])
        s.close()
    t.close()


def test_other_world_boxes(oracle_lib, beam_lib, cuda_device):
    """bmInsertMeshInTree takes the world box as an argument; non-default boxes must subdivide like the reference"""
    from raytracercuda_b200 import BeamError
    from raytracercuda_b200.tree import Tree
    rng = np.random.default_rng(5)
    P, I = random_soup(rng, 3000, 1.9, 0.04)
    for wmin, wmax in (((-2, -2, -2), (2, 2, 2)), ((-3, -1.5, -2), (3, 1.5, 2)), ((-7.5, -7.5, -7.5), (7.5, 7.5, 7.5))):
        parts = [(P, I)]
        t = _gpu_tree(cuda_device, parts, world_min=wmin, world_max=wmax)
        s = _oracle(oracle_lib, parts, wmin=wmin, wmax=wmax)
        assert t.stats()["depth"] == s.stats()["max_depth"] == s.stats()["min_leaf_depth"]
        _check_tree_against_oracle(t, s, parts)
        t.close(); s.close()
    with pytest.raises(BeamError):          # +-1000 needs more than 32 code bits
        Tree(cuda_device, world_min=(-1000, -1000, -1000), world_max=(1000, 1000, 1000))


def test_large_soup_properties(beam_lib, cuda_device):
    """size-independent properties at a size the oracle would need minutes for (2M triangles): sorted, every pair's code is
    a cell its triangle's box overlaps, leaves partition the pair list, tree is a proper binary tree"""
    from raytracercuda_b200 import synth
    from raytracercuda_b200.tree import Tree
    P, I = synth.soup(2_000_000)
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    st = t.stats()
    code, tri = t.pairs()
    assert code.size == st["pairs"] and st["nodes"] == st["leaves"] - 1
    key = (code.astype(np.uint64) << np.uint64(32)) | tri.astype(np.uint64)
    assert (key[1:] > key[:-1]).all()
    assert np.unique(code).size == st["leaves"]
    assert np.array_equal(np.unique(tri), np.arange(2_000_000, dtype=np.uint32))       # every in-world triangle lands somewhere
    # decode the x-first interleaved 31-bit path code and check the cell against the triangle's bounding box
    def compact(v):
        v = v & 0x09249249; v = (v | (v >> 2)) & 0x030C30C3; v = (v | (v >> 4)) & 0x0300F00F; v = (v | (v >> 8)) & 0x030000FF; v = (v | (v >> 16)) & 0x3FF
        return v
    c = code.astype(np.uint64) >> np.uint64(1)
    ix = (compact((c >> np.uint64(2)).astype(np.uint32)).astype(np.int64) << 1) | (code & 1).astype(np.int64)
    iy = compact((c >> np.uint64(1)).astype(np.uint32)).astype(np.int64); iz = compact(c.astype(np.uint32)).astype(np.int64)
    T = P.reshape(-1, 3, 3)[tri]
    for idx, w, a in ((ix, 60.0 / 2048, 0), (iy, 60.0 / 1024, 1), (iz, 60.0 / 1024, 2)):
        lo = -30.0 + idx * w
        assert (T[:, :, a].min(1) <= lo + w + 1e-4).all() and (T[:, :, a].max(1) >= lo - 1e-4).all()
    plen, prefix, c0, c1 = t.topology()
    refs = np.concatenate([c0, c1])
    assert np.array_equal(np.sort(refs[refs >= 0x80000000] & 0x7FFFFFFF), np.arange(st["leaves"]))
    t.close()


def test_flat_keygen_equals_descent_keygen(beam_lib, cuda_device, monkeypatch):
    """The fast key generator (per-candidate test with the margin-guarded ancestor walk) against the per-triangle reference descent,
    on the GPU, at a size the CPU oracle cannot reach: 3M triangles of mixed sizes, many hugging the big split planes."""
    from raytracercuda_b200.tree import Tree
    rng = np.random.default_rng(33)
    n = 3_000_000
    c = rng.uniform(-29.5, 29.5, (n, 1, 3)).astype(np.float32)
    snap = rng.integers(0, 4, n) == 0
    planes = np.array([0.0, 15.0, -15.0, 7.5, -7.5, 3.75, -3.75, 1.875], np.float32)
    ax = rng.integers(0, 3, n)
    c[snap, 0, ax[snap]] = planes[rng.integers(0, planes.size, snap.sum())] + rng.choice(np.array([0, 1e-7, -1e-7, 1e-5, -1e-4, 1e-3], np.float32), snap.sum())
    size = rng.choice(np.array([0.003, 0.01, 0.03, 0.1], np.float32), (n, 1, 1))
    P = (c + rng.uniform(-1, 1, (n, 3, 3)).astype(np.float32) * size).reshape(-1, 3)
    I = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
    out = {}
    for mode in ("flat", "descent"):
        monkeypatch.setenv("BM_KEYGEN", mode)
        t = Tree(cuda_device); t.add_mesh(P, I); t.build()
        out[mode] = (t.stats(), *t.pairs())
        t.close()
    assert out["flat"][0] == out["descent"][0]
    assert np.array_equal(out["flat"][1], out["descent"][1]) and np.array_equal(out["flat"][2], out["descent"][2])
    assert out["flat"][0]["pairs"] > 2 * n


def test_three_kernel_sort_path_and_sub_key_widths_give_the_same_tree(meshes, beam_lib, cuda_device):
    """The one-sweep radix passes against the three-kernel passes kept for >= 2^30 pairs (BM_SORT=3k), and the coarse against the fine
    in-cell sub-key (BM_SUB_BITS): the canonical pair list, the topology and the statistics may not depend on either.  The knobs are read
    once per process, so every variant runs in its own interpreter."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    prog = (
        "import sys, json, zlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from raytracercuda_b200.tree import Tree\n"
        "z = np.load(%r)\n"
        "out = {}\n"
        "for name in ('bunny', 'f16'):\n"
        "    t = Tree('cuda:0'); t.add_mesh(z[name + '_pos'], z[name + '_idx']); t.build()\n"
        "    c, tri = t.pairs(); topo = t.topology()\n"
        "    out[name] = [t.stats(), zlib.crc32(c.tobytes()), zlib.crc32(tri.tobytes())] + [zlib.crc32(np.sort(np.asarray(a)).tobytes()) for a in topo[:2]]\n"
        "    t.close()\n"
        "print(json.dumps(out))\n") % (ROOT, os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    res = {}
    for tag, env in (("default", {}), ("3k", {"BM_SORT": "3k"}), ("fine", {"BM_SUB_BITS": "15"}), ("coarse3k", {"BM_SUB_BITS": "3", "BM_SORT": "3k"})):
        r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=300, env={**os.environ, **env})
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    for tag in ("3k", "fine", "coarse3k"):
        assert res[tag] == res["default"], tag
    assert res["default"]["bunny"][0]["pairs"] == 184931
