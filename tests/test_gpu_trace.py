"""GPU tier: closest-hit traversal through the C ABI against the oracle on identical ray buffers.
Bar: reference-mode triangle id equal on >= 99.99 % of rays (ties in t excluded), t within 1e-5 relative."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_soup
from oracle.bindings import OracleScene

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1
EYE = [0.0, 0.0, -2.1]
ID3 = np.eye(3, dtype=np.float32)


def _compare(h, r, bases, min_agree=0.9999, what=""):
    """h: GPU hits dict; r: oracle dict.  Returns the agreement fraction."""
    otri = np.where(r["tri"] == 0xFFFFFFFF, np.uint32(0xFFFFFFFF), bases[np.minimum(r["mesh"], len(bases) - 1)] + r["tri"]).astype(np.uint32)
    considered = r["tie"] == 0
    same = (h["tri"] == otri)
    agree = same[considered].mean()
    assert agree >= min_agree, f"{what}: triangle id agreement {agree:.6f}"
    both = same & (otri != 0xFFFFFFFF)
    rel = np.abs(h["t"][both] - r["t"][both]) / np.maximum(np.abs(r["t"][both]), 1e-20)
    assert rel.size == 0 or rel.max() <= 1e-5, f"{what}: t differs by {rel.max():.3g} relative"
    miss = same & (otri == 0xFFFFFFFF)
    assert (h["t"][miss] == np.finfo(np.float32).max).all()
    return agree, (rel.max() if rel.size else 0.0)


@pytest.mark.parametrize("name,W,H", [("f16", 500, 500), ("suzanne", 320, 180), ("bunny", 640, 360)])
def test_camera_reference_mode(name, W, H, meshes, golden, oracle_lib, beam_lib, cuda_device):
    import torch
    from raytracercuda_b200.tree import Tree, hits_to_numpy, make_mesh_table
    P, I, N = meshes[name]
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    s = OracleScene(oracle_lib); s.add_mesh(P, I, N); s.build()
    rays = oracle_lib.initial_rays(W, H)
    r = s.trace_camera(rays, EYE, ID3, 0, NT)
    Pd, Id = t._meshes[0]
    Nd = torch.from_numpy(N).to(cuda_device)
    table = make_mesh_table([(Pd, Id, Nd)], cuda_device)
    px = torch.empty(W * H, dtype=torch.int32, device=cuda_device)
    hits, cnt = t.trace_camera(rays, EYE, ID3, pixels=px, mesh_table=table, counters=True, width=W)      # 8x4 tile order
    h = hits_to_numpy(hits)
    agree, rel = _compare(h, r, t.mesh_bases(), what=name)
    assert agree == 1.0 and rel == 0.0, "bundled meshes are expected to match the oracle bit for bit"
    assert np.array_equal(h["u"][h["tri"] != 0xFFFFFFFF], r["u"][r["tri"] != 0xFFFFFFFF])
    gp = px.cpu().numpy().view(np.uint32)
    assert np.array_equal(gp, r["pixels"]), "shaded pixels differ"
    res = "%dx%d" % (W, H)
    if res in golden["meshes"][name]["trace"]:
        assert int((gp != 255 << 8).sum()) == golden["meshes"][name]["trace"][res]["hits"]
    c = cnt.cpu().numpy()
    assert 0 < c[0] / (W * H) < r["node_visits"] / (W * H)        # fewer node visits than the reference's chain walk
    t.close(); s.close()


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_reference_pixels_golden(name, meshes, beam_lib, cuda_device):
    """against the unmodified reference's own pixels (tests/golden/pixels_*.npz)"""
    import torch
    from raytracercuda_b200.tree import Tree, make_mesh_table
    P, I, N = meshes[name]
    want = np.load(os.path.join(GOLDEN, "pixels_%s_200x200.npz" % name))["pixels"].reshape(-1)
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    Pd, Id = t._meshes[0]
    table = make_mesh_table([(Pd, Id, torch.from_numpy(N).to(cuda_device))], cuda_device)
    from raytracercuda_b200.host import ICamera
    cam = ICamera.create(); assert cam.setInitialRays(200, 200, -1, 1, -1, 1, 1) == 0
    px = torch.empty(200 * 200, dtype=torch.int32, device=cuda_device)
    t.trace_camera(cam.initialRays(), EYE, ID3, want_hits=False, pixels=px, mesh_table=table)
    assert np.array_equal(px.cpu().numpy().view(np.uint32), want)
    t.close()


def _ray_batch(rng, n, ext):
    o = rng.uniform(-ext, ext, (n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    R = np.concatenate([o, d], 1).astype(np.float32)
    R[::97, 3] = 0.0          # axis-parallel components: inf / NaN slab paths
    R[::193, 4] = 0.0
    return R


@pytest.mark.parametrize("mode", [0, 1])
def test_incoherent_rays_on_meshes(mode, meshes, oracle_lib, beam_lib, cuda_device):
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    rng = np.random.default_rng(11)
    for name in ("suzanne", "bunny"):
        P, I, N = meshes[name]
        t = Tree(cuda_device); t.add_mesh(P, I); t.build()
        s = OracleScene(oracle_lib); s.add_mesh(P, I, N); s.build()
        R = _ray_batch(rng, 60000, 2.5)          # origins inside and outside the model, incl. inside populated cells
        r = s.trace_rays(R[:, :3], R[:, 3:], mode, NT)
        h = hits_to_numpy(t.trace_rays(R, mode))
        agree, rel = _compare(h, r, t.mesh_bases(), what=f"{name} mode {mode}")
        if mode == 0:
            assert (r["t"][r["tri"] != 0xFFFFFFFF] < 0).any(), "the batch should exercise the reference's negative-t hits"
        t.close(); s.close()


def test_incoherent_rays_on_soup(oracle_lib, beam_lib, cuda_device):
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    rng = np.random.default_rng(12)
    P, I = random_soup(rng, 30000, 2.0, 0.15)
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
    R = _ray_batch(rng, 50000, 2.2)
    for mode in (0, 1):
        r = s.trace_rays(R[:, :3], R[:, 3:], mode, NT)
        h = hits_to_numpy(t.trace_rays(R, mode))
        _compare(h, r, t.mesh_bases(), what=f"soup mode {mode}")
        assert (r["tri"] != 0xFFFFFFFF).mean() > 0.3
    t.close(); s.close()


def test_reference_vs_exhaustive_closest(meshes, oracle_lib, beam_lib, cuda_device):
    """report-style check: on the bundled meshes the reference's first-leaf answer is the true closest hit for >= 99.99 %"""
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    for name in ("suzanne", "bunny"):
        P, I, N = meshes[name]
        t = Tree(cuda_device); t.add_mesh(P, I); t.build()
        rays = oracle_lib.initial_rays(400, 400)
        h0 = hits_to_numpy(t.trace_camera(rays, EYE, ID3, mode=0)[0]); h1 = hits_to_numpy(t.trace_camera(rays, EYE, ID3, mode=1)[0])
        assert (h0["tri"] == h1["tri"]).mean() >= 0.9999
        t.close()


def test_edge_cases(oracle_lib, beam_lib, cuda_device):
    import torch
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    rays = oracle_lib.initial_rays(16, 16)
    # empty scene
    t = Tree(cuda_device); t.build()
    h = hits_to_numpy(t.trace_camera(rays, EYE, ID3)[0])
    assert (h["tri"] == 0xFFFFFFFF).all() and (h["t"] == np.finfo(np.float32).max).all()
    assert t.trace_rays(np.zeros((0, 6), np.float32)).shape[0] == 0
    t.close()
    # single leaf (root is a leaf) and the viewer's commented two-triangle scene (Program.cpp:153-185)
    P = np.array([[-1, -1, 1.56], [0, 1, 1.56], [1, -1, 1.56], [2, 1, 1.56]], np.float32)
    for I in (np.array([[0, 1, 2]], np.uint32), np.array([[0, 1, 2], [1, 2, 3]], np.uint32)):
        t = Tree(cuda_device); t.add_mesh(P, I); t.build()
        s = OracleScene(oracle_lib); s.add_mesh(P, I); s.build()
        r = s.trace_camera(oracle_lib.initial_rays(64, 64), EYE, ID3, 0)
        h = hits_to_numpy(t.trace_camera(oracle_lib.initial_rays(64, 64), EYE, ID3)[0])
        _compare(h, r, t.mesh_bases(), min_agree=1.0)
        assert (h["tri"] != 0xFFFFFFFF).sum() > 100
        t.close(); s.close()
    # leaf cap: only the first 256 faces of a leaf are tested in reference mode
    tri1 = np.array([[0.001, 0.001, 0.5], [0.02, 0.001, 0.5], [0.001, 0.02, 0.5]], np.float32)
    Pc = np.tile(tri1, (300, 1)); Pc[3 * 299:, 2] = 0.49      # the 300th copy is nearer to the eye but beyond the cap
    Ic = np.arange(900, dtype=np.uint32).reshape(-1, 3)
    t = Tree(cuda_device); t.add_mesh(Pc, Ic); t.build()
    s = OracleScene(oracle_lib); s.add_mesh(Pc, Ic); s.build()
    R = np.array([[0.005, 0.005, -1, 0, 0, 1]], np.float32)
    r = s.trace_rays(R[:, :3], R[:, 3:], 0)
    h = hits_to_numpy(t.trace_rays(R, 0))
    assert h["tri"][0] == r["tri"][0] == 0 and h["t"][0] == r["t"][0]
    h1 = hits_to_numpy(t.trace_rays(R, 1))
    assert h1["tri"][0] == 299
    t.close(); s.close()


def test_bounce_rays_workflow(oracle_lib, beam_lib, cuda_device):
    """config-2 style frame on a small stand-in: primary -> diffuse bounce rays (GPU-generated) -> bounce trace with the count kept
    on the device; the oracle traces the SAME bounce buffer (parity on identical ray buffers)."""
    import ctypes as C
    import torch
    from raytracercuda_b200 import synth, workloads
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    wl = workloads.get("c2small")
    W, H = wl.width, wl.height
    n = W * H
    t = Tree(cuda_device); t.add_mesh(wl.positions, wl.indices); t.build()
    s = OracleScene(oracle_lib); s.add_mesh(wl.positions, wl.indices, wl.normals); s.build()
    rays = oracle_lib.initial_rays(W, H)
    rd = torch.from_numpy(rays).to(cuda_device)
    hits1, _ = t.trace_camera(rd, synth.EYE, synth.IDENTITY, width=W)
    f3 = (C.c_float * 3)(*synth.EYE); f9 = (C.c_float * 9)(*synth.IDENTITY.reshape(-1))
    out = torch.zeros((n, 6), dtype=torch.float32, device=cuda_device); src = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    cnt = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    assert beam_lib.bmBounceRaysCamera(t.h, rd.data_ptr(), f3, f9, hits1.data_ptr(), n, 0, 1, 7, out.data_ptr(), src.data_ptr(), cnt.data_ptr()) == 0
    nb = int(cnt.cpu()[0])
    h1 = hits_to_numpy(hits1)
    hit_idx = np.flatnonzero(h1["tri"] != 0xFFFFFFFF)
    assert nb == hit_idx.size > 1000
    assert np.array_equal(src.cpu().numpy()[:nb], hit_idx)                       # compacted in ray order
    # same rays through the explicit-ray entry point (bmCameraRays + bmBounceRays)
    cam_rays = t.camera_rays(rd, synth.EYE, synth.IDENTITY, width=W)
    out2, src2, cnt2 = t.bounce_rays(cam_rays, hits1, sample=0, bounce=1, seed=7)
    assert int(cnt2.cpu()[0]) == nb and torch.equal(out2[:nb], out[:nb])
    B = out[:nb].cpu().numpy()
    # geometry of the construction: origin sits 1e-4 off the hit point on the side the ray came from, direction is a unit vector
    # in the hemisphere of that normal
    o1 = np.asarray(synth.EYE, np.float32) + h1["t"][hit_idx, None] * rays[hit_idx]
    off = B[:, :3] - o1
    assert np.abs(np.linalg.norm(off, axis=1) - 1e-4).max() < 2e-6
    assert np.abs(np.linalg.norm(B[:, 3:], axis=1) - 1).max() < 1e-5
    assert ((off * B[:, 3:]).sum(1) > -1e-9).all() and ((off * rays[hit_idx]).sum(1) < 0).all()
    # bounce trace with the device-side count vs the oracle on the same buffer
    hits2 = torch.full((n, 4), -1.0, dtype=torch.float32, device=cuda_device)
    assert beam_lib.bmTraceRaysIndirect(t.h, out.data_ptr(), cnt.data_ptr(), n, 0, hits2.data_ptr(), None) == 0
    h2 = hits_to_numpy(hits2[:nb])
    assert (hits2[nb:].cpu().numpy() == -1.0).all()                               # nothing beyond the device count is touched
    r = s.trace_rays(B[:, :3], B[:, 3:], 0, NT)
    _compare(h2, r, t.mesh_bases(), what="bounce rays")
    # In the reference's any-sign semantics nearly every bounce ray "hits" its own source triangle behind its origin (t = -1e-4 / cos):
    # that is why the product traces bounce rays as closest hits with t >= 0 (host.cu marchProgressive, bench.py) -- same check there
    self_hit = (h2["tri"] == h1["tri"][hit_idx]) & (h2["t"] < 0)
    assert self_hit.mean() > 0.9
    assert beam_lib.bmTraceRaysIndirect(t.h, out.data_ptr(), cnt.data_ptr(), n, 1, hits2.data_ptr(), None) == 0
    h3 = hits_to_numpy(hits2[:nb])
    r3 = s.trace_rays(B[:, :3], B[:, 3:], 1, NT)
    _compare(h3, r3, t.mesh_bases(), what="bounce rays, closest mode")
    assert (h3["t"][h3["tri"] != 0xFFFFFFFF] >= 0).all() and (h3["tri"] == 0xFFFFFFFF).mean() > 0.5
    t.close(); s.close()


@pytest.mark.parametrize("layout", ["binary", "wide"])
def test_export_import_replica_traces_identically(layout, meshes, beam_lib, cuda_device, monkeypatch):
    """multi-GPU replication path on one GPU: export the traversal arrays, adopt them as a read-only tree, trace both -- for both node
    layouts (the wide records are decoded against the source tree's grid, which travels with the import call)"""
    import ctypes as C
    import torch
    from raytracercuda_b200 import _lib
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    monkeypatch.setenv("BM_NODE_LAYOUT", layout)
    P, I, N = meshes["bunny"]
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    assert t.node_layout() == (_lib.BM_NODES_WIDE if layout == "wide" else _lib.BM_NODES_BINARY)
    sz = (C.c_uint64 * 8)()
    assert beam_lib.bmTreeArraySizes(t.h, sz) == 0
    st = t.stats()
    assert (sz[5], sz[6], sz[7]) == (st["leaves"], st["pairs"], st["triangles"])
    assert sz[0] == st["nodes"] * (64 if layout == "wide" else 128)
    bufs = [torch.empty(max(int(b), 16), dtype=torch.uint8, device=cuda_device) for b in sz[:5]]
    assert beam_lib.bmTreeExport(t.h, *[b.data_ptr() for b in bufs]) == 0
    torch.cuda.synchronize()
    h = C.c_void_p()
    wmin, wmax, min_leaf, max_depth = t.world()
    assert beam_lib.bmTreeImport(C.byref(h), *[b.data_ptr() for b in bufs], sz[5], sz[6], sz[7], t.root(), t.node_layout(), wmin.ctypes.data, wmax.ctypes.data,
                                 min_leaf, max_depth) == 0
    rng = np.random.default_rng(5)
    R = torch.from_numpy(_ray_batch(rng, 40000, 2.5)).to(cuda_device)
    for mode in (0, 1):
        a = hits_to_numpy(t.trace_rays(R, mode))
        hits = torch.empty((R.shape[0], 4), dtype=torch.float32, device=cuda_device)
        assert beam_lib.bmTraceRays(h, R.data_ptr(), R.shape[0], mode, hits.data_ptr(), None) == 0
        b = hits_to_numpy(hits)
        assert np.array_equal(a["tri"], b["tri"]) and np.array_equal(a["t"], b["t"])
    assert beam_lib.bmTreeBuild(h) != 0                      # read-only
    st8 = (C.c_uint64 * 8)()
    assert beam_lib.bmTreeStats(h, st8) == 0 and st8[2] == st["leaves"]
    beam_lib.bmTreeDestroy(h)
    t.close()


def test_wide_records_trace_like_binary_records_and_enclose_the_exact_boxes(meshes, oracle_lib, beam_lib, cuda_device, monkeypatch):
    """The two node layouts of one tree (BM_NODE_LAYOUT=both): identical hits in both trace modes on camera rays and incoherent batches,
    fewer node visits through the wide records, and every decoded (quantised) slot box encloses the exact box of the child it stands for."""
    import torch
    from raytracercuda_b200 import _lib, synth
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    monkeypatch.setenv("BM_NODE_LAYOUT", "both")
    rng = np.random.default_rng(11)
    Ps, Is = random_soup(rng, 150000, 4.0, 0.05)
    for name, (P, I) in {"bunny": meshes["bunny"][:2], "f16": meshes["f16"][:2], "soup": (Ps, Is)}.items():
        t = Tree(cuda_device); t.add_mesh(P, I); t.build()
        assert t.node_layout() == _lib.BM_NODES_BOTH
        # slot boxes vs the exact child boxes: slots (0,1) are the children of child 0 when it is internal, else slot 0 is child 0 itself
        plen, prefix, c0, c1 = t.topology()
        exact = t.child_boxes().reshape(-1, 2, 6)                         # per node: boxes of its two children
        wb, wr = t.wide_slots()
        LEAF = 0x80000000
        for h, c in ((0, c0), (1, c1)):
            inner = (c & LEAF) == 0
            ci = c[inner].astype(np.int64)
            for k in (0, 1):                                              # absorbed child: its children's refs and boxes
                ck = (c0 if k == 0 else c1)[ci]
                assert np.array_equal(wr[inner, 2 * h + k], ck)
                ex = exact[ci, k]
                got = wb[inner, 2 * h + k]
                assert (got[:, :3] <= ex[:, :3]).all() and (got[:, 3:] >= ex[:, 3:]).all(), name
                slack = np.maximum(ex[:, :3] - got[:, :3], got[:, 3:] - ex[:, 3:]).max(1)
                size = (ex[:, 3:] - ex[:, :3]).max(1)
                assert np.median(slack / np.maximum(size, 1e-6)) < 0.05     # 8 bits per coordinate on the node's own grid
            leaf = ~inner
            assert np.array_equal(wr[leaf, 2 * h], c[leaf]) and (wr[leaf, 2 * h + 1] == 0x7FFFFFFF).all()
            got = wb[leaf, 2 * h]; ex = exact[leaf, h]
            assert (got[:, :3] <= ex[:, :3]).all() and (got[:, 3:] >= ex[:, 3:]).all()
        rays = _ray_batch(rng, 60000, 2.5 if name != "soup" else 4.5)
        R = torch.from_numpy(rays).to(cuda_device)
        cam = torch.from_numpy(oracle_lib.initial_rays(320, 200)).to(cuda_device)
        # results through a binary-only and a wide-only build of the same scene
        t.close()
        outs = {}
        for lay in ("binary", "wide"):
            monkeypatch.setenv("BM_NODE_LAYOUT", lay)
            tt = Tree(cuda_device); tt.add_mesh(P, I); tt.build()
            o = []
            for mode in (0, 1):
                hits, cnt = tt.trace_rays(R, mode, counters=True)
                o.append((hits_to_numpy(hits), int(cnt.cpu()[0])))
                hits, cnt = tt.trace_camera(cam, [0, 0, -2.1] if name != "soup" else [0, 0, -6.0], ID3, mode=mode, counters=True, width=320)
                o.append((hits_to_numpy(hits), int(cnt.cpu()[0])))
            outs[lay] = o
            tt.close()
        monkeypatch.setenv("BM_NODE_LAYOUT", "both")
        for k, ((hb, vb), (hw, vw)) in enumerate(zip(outs["binary"], outs["wide"])):
            assert np.array_equal(hb["tri"], hw["tri"]) and np.array_equal(hb["t"], hw["t"]) and np.array_equal(hb["u"], hw["u"]), name
            if k < 2:        # reference mode; in closest mode the binary walk also cuts every child box down to its half of the node's cell
                assert vw < vb, "the wide walk must take fewer node visits"


@pytest.mark.parametrize("name,W,H", [("bunny", 640, 360), ("suzanne", 203, 77)])
def test_repeated_frames_follow_the_tile_schedule_and_do_not_change(name, W, H, meshes, oracle_lib, beam_lib, cuda_device):
    """From the second camera launch of a frame size on, the kernel follows the work list built from the previous launch's tile costs
    (expensive tiles first, the costliest as four quarter entries).  Every frame -- same view, a moved eye (stale costs), another
    resolution in between (state reset) -- must equal the oracle / the first frame bit for bit; partial tiles included (203x77)."""
    import torch
    from raytracercuda_b200.tree import Tree, hits_to_numpy, make_mesh_table
    P, I, N = meshes[name]
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    s = OracleScene(oracle_lib); s.add_mesh(P, I, N); s.build()
    Pd, Id = t._meshes[0]
    Nd = torch.from_numpy(N).to(cuda_device)
    table = make_mesh_table([(Pd, Id, Nd)], cuda_device)
    rays = oracle_lib.initial_rays(W, H)
    px = torch.empty(W * H, dtype=torch.int32, device=cuda_device)
    want = {}
    for eye in (EYE, [0.3, 0.1, -2.0]):
        r = s.trace_camera(rays, eye, ID3, 0, NT)
        otri = np.where(r["tri"] == 0xFFFFFFFF, np.uint32(0xFFFFFFFF), r["tri"]).astype(np.uint32)
        want[tuple(eye)] = (otri, r["t"], r["pixels"], r["tie"] == 0)
    seq = [EYE, EYE, EYE, [0.3, 0.1, -2.0], [0.3, 0.1, -2.0], EYE, EYE]
    first = None
    for k, eye in enumerate(seq):
        hits, _ = t.trace_camera(rays, eye, ID3, pixels=px, mesh_table=table, width=W)
        h = hits_to_numpy(hits); gp = px.cpu().numpy().view(np.uint32).copy()
        otri, ot, opx, ok = want[tuple(eye)]
        assert np.array_equal(h["tri"][ok], otri[ok]), "frame %d" % k
        both = ok & (otri != 0xFFFFFFFF)
        assert np.array_equal(h["t"][both], ot[both]), "frame %d" % k
        assert np.array_equal(gp[ok], opx[ok]), "frame %d" % k
        if k == 0:
            first = (h["tri"].copy(), h["t"].copy(), gp)
        if k in (1, 2, 5, 6):
            assert np.array_equal(h["tri"], first[0]) and np.array_equal(h["t"], first[1]) and np.array_equal(gp, first[2]), "frame %d differs from frame 0" % k
        if k == 4:
            # another frame size in between resets the schedule
            r2 = oracle_lib.initial_rays(64, 40)
            h2, _ = t.trace_camera(r2, EYE, ID3, width=64)
            o2 = s.trace_camera(r2, EYE, ID3, 0, NT)
            ok2 = o2["tie"] == 0
            assert np.array_equal(hits_to_numpy(h2)["tri"][ok2], o2["tri"][ok2])
    t.close(); s.close()


def test_tree_views_trace_concurrently_and_identically(meshes, beam_lib, cuda_device):
    """bmTreeView: two more handles on one built tree, each on its own stream, tracing the two halves of a frame at the same time;
    the halves must equal the single-handle frame, and a view of an unbuilt tree is refused."""
    import ctypes as C
    import torch
    from raytracercuda_b200 import _lib
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    P, I, N = meshes["bunny"]
    t = Tree(cuda_device)
    h = C.c_void_p()
    assert beam_lib.bmTreeView(t.h, C.byref(h)) == 102                      # BM_ERROR_NOT_BUILT
    t.add_mesh(P, I); t.build()
    W, H = 320, 200
    from raytracercuda_b200.host import ICamera
    cam = ICamera.create(); cam.setInitialRays(W, H, -1, 1, -1, 1, 1)
    rays = torch.from_numpy(cam.initialRays()).to(cuda_device)
    whole, _ = t.trace_camera(rays, EYE, ID3, width=W)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(cuda_device), torch.cuda.Stream(cuda_device)
    v1, v2 = t.view(s1), t.view(s2)
    half = (H // 2) * W
    for rep in range(3):                                                     # repeated: the views keep their own tile schedules
        a, _ = v1.trace_camera(rays[:half], EYE, ID3, width=W)
        b, _ = v2.trace_camera(rays[half:], EYE, ID3, width=W)
        s1.synchronize(); s2.synchronize()
        got = torch.cat([a, b])
        assert torch.equal(got.view(torch.int32), whole.view(torch.int32)), "rep %d" % rep
    assert v1.stats()["leaves"] == t.stats()["leaves"]
    v1.close(); v2.close(); t.close()


def test_tiled_ray_batches_equal_plain_batches(meshes, beam_lib, cuda_device):
    """bmTraceRaysTiled (image-shaped batches in 8x4 tiles under the tile schedule) against bmTraceRays on the jittered camera rays of
    several samples, repeated so that the schedule of the previous launch is in use; partial tiles included (301 x 203)."""
    import ctypes as C
    import torch
    from raytracercuda_b200 import _lib
    from raytracercuda_b200.host import ICamera
    from raytracercuda_b200.tree import Tree
    L = beam_lib
    P, I, N = meshes["bunny"]
    t = Tree(cuda_device); t.add_mesh(P, I); t.build()
    W, H = 301, 203
    n = W * H
    cam = ICamera.create(); cam.setInitialRays(W, H, -1, 1, -1, 1, 1)
    table = torch.from_numpy(cam.initialRays()).to(cuda_device)
    a = torch.empty((n, 4), dtype=torch.float32, device=cuda_device); b = torch.empty_like(a)
    for sample in range(5):
        rays = t.camera_rays(table, EYE, ID3, width=W, jitter=(2.0 / W, 2.0 / H), sample=sample, seed=3)
        for mode in (0, 1):
            _lib.check(L.bmTraceRays(t.h, rays.data_ptr(), n, mode, a.data_ptr(), None), "bmTraceRays")
            _lib.check(L.bmTraceRaysTiled(t.h, rays.data_ptr(), n, W, mode, b.data_ptr(), None), "bmTraceRaysTiled")
            torch.cuda.synchronize()
            assert torch.equal(a.view(torch.int32), b.view(torch.int32)), "sample %d mode %d" % (sample, mode)
    assert (a[:, 3].view(torch.int32) != -1).sum() > 5000
    cnt = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    _lib.check(L.bmTraceRaysTiled(t.h, rays.data_ptr(), n, W, 0, b.data_ptr(), cnt.data_ptr()), "bmTraceRaysTiled")
    assert int(cnt.cpu()[0]) > n
    assert L.bmTraceRaysTiled(t.h, rays.data_ptr(), n, 0, 0, b.data_ptr(), None) == 2 and L.bmTraceRaysTiled(t.h, rays.data_ptr(), n, 11, 0, b.data_ptr(), None) == 2      # 301 * 203 is not a multiple of 11
    t.close()


@pytest.mark.parametrize("layout", ["binary", "wide"])
def test_bricked_ray_batches_equal_the_direct_traversal(layout, meshes, oracle_lib, beam_lib, cuda_device, monkeypatch):
    """bmTreeSetBricks: a batch cut into (ray, brick) segments, sorted by brick and merged per ray must give the hits of the direct
    traversal bit for bit -- both modes, both node layouts, several brick depths (also depths that are no multiple of 3 and bricks smaller
    than the populated region), rays with axis-parallel components, origins inside populated cells, a replica made by export/import --
    and agree with the oracle like the direct traversal does."""
    import ctypes as C
    import torch
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    monkeypatch.setenv("BM_NODE_LAYOUT", layout)
    rng = np.random.default_rng(21)
    Ps, Is = random_soup(rng, 40000, 6.0, 0.08)
    scenes = [("soup", Ps, Is, None, 6.5), ("bunny", *meshes["bunny"], 2.5)]
    for name, P, I, N, ext in scenes:
        t = Tree(cuda_device); t.add_mesh(P, I); t.build()
        R = _ray_batch(rng, 120000, ext)
        R[::7, :3] *= 3.0                                  # origins far outside the populated region too
        Rd = torch.from_numpy(R).to(cuda_device)
        s = OracleScene(oracle_lib); s.add_mesh(P, I, N); s.build(NT)
        for mode in (0, 1):
            t.set_bricks(0)
            a, ca = t.trace_rays(Rd, mode, counters=True)
            a = hits_to_numpy(a)
            r = s.trace_rays(R[:, :3], R[:, 3:], mode, NT)
            _compare(a, r, t.mesh_bases(), what=f"{name} direct mode {mode}")
            for bits in (3, 6, 10, 12, 15):
                t.set_bricks(1, bits)
                b, cb = t.trace_rays(Rd, mode, counters=True)
                b = hits_to_numpy(b)
                for k in ("tri", "t", "u", "v"):
                    assert np.array_equal(a[k], b[k]), f"{name} {layout} mode {mode} bits {bits}: {k} differs on {(a[k] != b[k]).sum()} rays"
                assert cb[0].item() > 0
        # a replica (imported arrays) finds its bricks by itself
        sz = (C.c_uint64 * 8)(); assert beam_lib.bmTreeArraySizes(t.h, sz) == 0
        bufs = [torch.empty(max(int(x), 16), dtype=torch.uint8, device=cuda_device) for x in sz[:5]]
        assert beam_lib.bmTreeExport(t.h, *[x.data_ptr() for x in bufs]) == 0
        torch.cuda.synchronize()
        h = C.c_void_p(); wmin, wmax, min_leaf, max_depth = t.world()
        assert beam_lib.bmTreeImport(C.byref(h), *[x.data_ptr() for x in bufs], sz[5], sz[6], sz[7], t.root(), t.node_layout(), wmin.ctypes.data, wmax.ctypes.data,
                                     min_leaf, max_depth) == 0
        assert beam_lib.bmTreeSetBricks(h, 1, 9) == 0
        hits = torch.empty((R.shape[0], 4), dtype=torch.float32, device=cuda_device)
        assert beam_lib.bmTraceRays(h, Rd.data_ptr(), R.shape[0], 1, hits.data_ptr(), None) == 0
        b = hits_to_numpy(hits)
        assert np.array_equal(a["tri"], b["tri"]) and np.array_equal(a["t"], b["t"])
        beam_lib.bmTreeDestroy(h)
        t.close(); s.close()
    assert beam_lib.bmTreeSetBricks(None, 0, 0) != 0


def test_bounce_launch_and_multi_sample_launch_equal_the_per_sample_lists(beam_lib, cuda_device):
    """bmTraceCameraSamples (several jittered samples of every pixel in one launch) and bmTraceBounceCamera (bounce rays formed in the
    kernel, no ray list) against the launches they replace: per sample bmTraceCameraJittered, bmBounceRaysCameraJittered ->
    bmTraceRays(closest) -> bmProgressiveAccumulate.  Same rays, same traversal: hit records bit-equal, accumulators equal."""
    import ctypes as C
    import torch
    from raytracercuda_b200 import host as H, workloads
    from raytracercuda_b200.tree import Tree
    L = beam_lib
    wl = workloads.get("c2small")
    W, Hh = 176, 100
    n = W * Hh
    cam = H.ICamera.create(); assert cam.setInitialRays(W, Hh, -1, 1, -1, 1, 1) == 0
    table = torch.from_numpy(cam.initialRays()).to(cuda_device)
    t = Tree(cuda_device); t.add_mesh(wl.positions, wl.indices); t.build()
    eye = (C.c_float * 3)(0.0, 0.0, -2.1); ori = (C.c_float * 9)(*np.eye(3, dtype=np.float32).reshape(-1))
    jx, jy = 2.0 / W, 2.0 / Hh
    SEED, S0, G = 7, 2, 3
    dev = cuda_device
    f4 = lambda m: torch.empty((m, 4), dtype=torch.float32, device=dev)
    for width in (W, ):
        # --- pixel-centre frame: bounce launch vs list ---------------------------------------------------------------------------
        h1 = f4(n)
        assert L.bmTraceCamera(t.h, table.data_ptr(), n, width, eye, ori, 0, h1.data_ptr(), None, None, None) == 0
        r2 = torch.empty((n, 6), dtype=torch.float32, device=dev); src = torch.empty(n, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        assert L.bmBounceRaysCamera(t.h, table.data_ptr(), eye, ori, h1.data_ptr(), n, 0, 1, SEED, r2.data_ptr(), src.data_ptr(), cnt.data_ptr()) == 0
        nb = int(cnt.item()); assert nb > 1000
        h2_list = f4(n)
        assert L.bmTraceRays(t.h, r2.data_ptr(), nb, 1, h2_list.data_ptr(), None) == 0
        for rep in range(2):                      # the second launch runs under the tile schedule the first one recorded
            h2 = torch.full((n, 4), -7.0, dtype=torch.float32, device=dev); count = torch.zeros(1, dtype=torch.int64, device=dev)
            acc = torch.zeros(n, dtype=torch.float32, device=dev); cov = torch.zeros(n, dtype=torch.int32, device=dev)
            assert L.bmTraceBounceCamera(t.h, table.data_ptr(), n, width, eye, ori, 0.0, 0.0, 0xFFFFFFFF, 1, 0, h1.data_ptr(), 0, 1, SEED, h2.data_ptr(), 0.25, 0.75,
                                         acc.data_ptr(), cov.data_ptr(), count.data_ptr(), None) == 0
            torch.cuda.synchronize()
            assert int(count.item()) == nb
            s_ = src[:nb].long()
            assert torch.equal(h2[s_].view(torch.int32), h2_list[:nb].view(torch.int32))
            hit = h1[:, 3].view(torch.int32) != -1
            assert bool((h2[~hit] == -7.0).all())                                           # written only where the primary ray hit
            want = torch.zeros(n, dtype=torch.float32, device=dev); want[hit] = 0.25
            esc = s_[h2_list[:nb, 3].view(torch.int32) == -1]; want[esc] += 0.75
            assert torch.equal(acc, want) and torch.equal(cov.bool(), hit)
        # --- G jittered samples in one launch vs one launch per sample -------------------------------------------------------------
        for rep in range(2):
            hs = f4(n * G)
            assert L.bmTraceCameraSamples(t.h, table.data_ptr(), n, width, eye, ori, jx, jy, S0, G, SEED, 0, hs.data_ptr(), None) == 0
            acc = torch.zeros(n, dtype=torch.float32, device=dev); cov = torch.zeros(n, dtype=torch.int32, device=dev); count = torch.zeros(1, dtype=torch.int64, device=dev)
            h2s = torch.full((n * G, 4), -7.0, dtype=torch.float32, device=dev)
            assert L.bmTraceBounceCamera(t.h, table.data_ptr(), n, width, eye, ori, jx, jy, S0, G, SEED, hs.data_ptr(), S0, 1, SEED, h2s.data_ptr(), 0.25, 0.75,
                                         acc.data_ptr(), cov.data_ptr(), count.data_ptr(), None) == 0
            want = torch.zeros(n, dtype=torch.float32, device=dev); wcov = torch.zeros(n, dtype=torch.bool, device=dev); total = 0
            for k in range(G):
                hk = f4(n)
                assert L.bmTraceCameraJittered(t.h, table.data_ptr(), n, width, eye, ori, jx, jy, S0 + k, SEED, 0, hk.data_ptr(), None, None, None) == 0
                assert torch.equal(hk.view(torch.int32), hs[k * n:(k + 1) * n].view(torch.int32)), "sample %d of the multi-sample launch differs" % k
                assert L.bmBounceRaysCameraJittered(t.h, table.data_ptr(), eye, ori, jx, jy, S0 + k, SEED, hk.data_ptr(), n, S0 + k, 1, SEED, r2.data_ptr(), src.data_ptr(), cnt.data_ptr()) == 0
                nbk = int(cnt.item()); total += nbk
                assert L.bmTraceRays(t.h, r2.data_ptr(), nbk, 1, h2_list.data_ptr(), None) == 0
                s_ = src[:nbk].long()
                assert torch.equal(h2s[k * n:(k + 1) * n][s_].view(torch.int32), h2_list[:nbk].view(torch.int32)), "bounce hits of sample %d differ" % k
                hit = hk[:, 3].view(torch.int32) != -1
                want[hit] += 0.25; wcov |= hit
                want[s_[h2_list[:nbk, 3].view(torch.int32) == -1]] += 0.75
            torch.cuda.synchronize()
            assert int(count.item()) == total
            assert torch.equal(acc, want) and torch.equal(cov.bool(), wcov)
    # argument checks: a pixel-centre frame is one sample; a multi-sample camera launch needs jittered samples
    assert L.bmTraceBounceCamera(t.h, table.data_ptr(), n, W, eye, ori, 0.0, 0.0, 0xFFFFFFFF, 2, 0, h1.data_ptr(), 0, 1, SEED, None, 0.25, 0.75, None, None, None, None) != 0
    assert L.bmTraceCameraSamples(t.h, table.data_ptr(), n, W, eye, ori, jx, jy, 0xFFFFFFFF, 1, SEED, 0, hs.data_ptr(), None) != 0
    t.close()
