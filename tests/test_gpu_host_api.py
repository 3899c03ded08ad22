"""GPU tier: the host-class mirror of the reference's public interface (Beam.h:32-72) end to end with HOST buffers, and
the reference's own extern "C" symbols (bmResetScene / bmInsertMeshInTree / bmMarch / bmClear)."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _scene(meshes, names):
    from raytracercuda_b200.host import IMesh, IScene, VERTEX_DATA_NORMAL, VERTEX_DATA_POSITION
    scene = IScene.create()
    for n in names:
        P, I, N = meshes[n]
        m = IMesh.create()
        assert m.setIndices(I, I.size) == 0
        assert m.setVertexData(P, P.shape[0], 3, VERTEX_DATA_POSITION) == 0
        assert m.setVertexData(N, N.shape[0], 3, VERTEX_DATA_NORMAL) == 0
        scene.addMesh(m)
    return scene


@pytest.mark.parametrize("name", ["f16", "suzanne", "bunny"])
def test_viewer_sequence_matches_reference_pixels(name, meshes, golden, beam_lib, cuda_device):
    """Program::initialize / renderFrame (TestProgram/Program.cpp:140-194, 287-355) replayed headless"""
    from raytracercuda_b200.host import ICamera, IRenderTarget
    scene = _scene(meshes, [name])
    scene.updateGPUScene()
    cam = ICamera.create()
    assert cam.setInitialRays(200, 200, -1, 1, -1, 1, 1) == 0
    assert zlib.crc32(cam.initialRays().tobytes()) == golden["initial_rays_200x200_crc32"]
    rt = IRenderTarget.createHeadless(200, 200)
    assert rt.lock() == 0
    assert cam.clear(0x12345678) == 0
    assert (rt.readPixels() == 0x12345678).all()
    assert cam.traceScene([0, 0, -2.1], np.eye(3, dtype=np.float32), scene) == 0
    px = rt.readPixels()
    want = np.load(os.path.join(GOLDEN, "pixels_%s_200x200.npz" % name))["pixels"]
    assert np.array_equal(px, want)
    err, hits = cam.traceSceneHits([0, 0, -2.1], np.eye(3, dtype=np.float32), scene)
    assert err == 0 and int((hits["tri"] != 0xFFFFFFFF).sum()) == golden["meshes"][name]["trace"]["200x200"]["hits"]
    assert rt.unlock() == 0


def test_error_codes(meshes, beam_lib, cuda_device):
    from raytracercuda_b200 import host as H
    m = H.IMesh.create()
    P, I, N = meshes["f16"]
    assert m.setVertexData(None, 10, 3, 0) == H.ERROR_INVALID_PARAMETER
    assert m.setVertexData(P, 0, 3, 0) == H.ERROR_INVALID_PARAMETER
    assert m.setVertexData(P, P.shape[0], 4, H.VERTEX_DATA_POSITION) == H.ERROR_INVALID_PARAMETER      # positions need 3 components
    assert m.setVertexData(P, P.shape[0], 5, H.VERTEX_DATA_EXTRA1) == H.ERROR_INVALID_PARAMETER
    assert m.setVertexData(P, P.shape[0], 3, H.VERTEX_DATA_COUNT) == H.ERROR_INVALID_PARAMETER
    assert m.setVertexData(P, P.shape[0], 3, H.VERTEX_DATA_POSITION) == 0
    assert m.setVertexData(N[:-1], P.shape[0] - 1, 3, H.VERTEX_DATA_NORMAL) == H.ERROR_INVALID_PARAMETER  # all slots share one vertex count
    assert m.setIndices(I, I.size - 1) == H.ERROR_INVALID_PARAMETER
    assert m.setIndices(None, 3) == H.ERROR_INVALID_PARAMETER
    assert m.setIndices(I, I.size) == 0
    cam = H.ICamera.create()
    assert cam.setInitialRays(0, 10) == H.ERROR_INVALID_PARAMETER
    scene = H.IScene.create(); scene.addMesh(m); scene.updateGPUScene()
    assert cam.traceScene([0, 0, -2.1], np.eye(3), scene) == H.ERROR_INVALID_PARAMETER                  # no rays yet
    assert cam.setInitialRays(64, 48, -1, 1, -1, 1, 1) == 0
    assert cam.clear(0) == H.ERROR_NO_RENDER_TARGET
    assert cam.traceScene([0, 0, -2.1], np.eye(3), scene) == H.ERROR_NO_RENDER_TARGET
    assert cam.traceScene(None, np.eye(3), scene) == H.ERROR_INVALID_PARAMETER
    assert H.IRenderTarget.createHeadless(0, 10) is None and H.IRenderTarget.createHeadless(64, 48, 64) is None
    rt = H.IRenderTarget.createHeadless(32, 32)
    assert rt.unlock() == H.ERROR_LOCK_FIRST
    assert rt.lock() == 0 and rt.lock() == H.ERROR_UNLOCK_FIRST
    assert cam.traceScene([0, 0, -2.1], np.eye(3), scene) == H.ERROR_RT_CAM_MISMATCH
    assert rt.unlock() == 0
    rt2 = H.IRenderTarget.createHeadless(64, 48)
    assert rt2.lock() == 0 and cam.traceScene([0, 0, -2.1], np.eye(3), scene) == 0
    assert (rt2.readPixels() != 255 << 8).sum() > 0
    # removeMesh + rebuild -> empty scene -> all miss
    scene.removeMesh(m); scene.updateGPUScene()
    assert cam.traceScene([0, 0, -2.1], np.eye(3), scene) == 0
    assert (rt2.readPixels() == 255 << 8).all()
    assert rt2.unlock() == 0


def test_reference_symbols_direct(meshes, golden, beam_lib, cuda_device):
    """SceneTree.cpp's own call sequence on the reference's symbols, with torch-owned device memory and lazy build in bmMarch"""
    import torch
    from raytracercuda_b200 import _lib
    from raytracercuda_b200.tree import make_mesh_table
    L = beam_lib
    P, I, N = meshes["suzanne"]
    dev = cuda_device
    Pd = torch.from_numpy(P).to(dev); Id = torch.from_numpy(I.view(np.int32)).to(dev).reshape(-1); Nd = torch.from_numpy(N).to(dev)
    root = torch.zeros(L.bmGetNodeSize(), dtype=torch.uint8, device=dev)
    assert [L.bmGetFaceSize(), L.bmGetFacePtrSize(), L.bmGetMaterialSize(), L.bmGetNodeSize(), L.bmGetFaceStoreSize(), L.bmGetNodeStoreSize(),
            L.bmGetFaceGroupStoreSize()] == [golden["struct_sizes"][k] for k in ("bmFace", "bmFacePtr", "bmMaterial", "bmTreeNode", "bmFaceStore", "bmNodeStore", "bmFaceGroupStore")]
    assert C.sizeof(_lib.bmStaticMeshData) == golden["struct_sizes"]["StaticMeshData"] == 136
    L.bmResetScene(root.data_ptr(), None, None, None, None, None, None, 1 << 24, 1 << 24, 1 << 24)
    lo, hi = _lib.bmVec3(-30, -30, -30), _lib.bmVec3(30, 30, 30)
    L.bmInsertMeshInTree(Pd.data_ptr(), Id.data_ptr(), Id.numel(), 0, C.byref(lo), C.byref(hi), root.data_ptr(), None, None, None, None)
    from raytracercuda_b200.host import ICamera
    cam = ICamera.create(); cam.setInitialRays(200, 200, -1, 1, -1, 1, 1)
    rays = torch.from_numpy(cam.initialRays()).to(dev)
    table = make_mesh_table([(Pd, Id, Nd)], dev)
    buf = torch.zeros(200 * 200, dtype=torch.int32, device=dev)
    eye = _lib.bmVec3(0, 0, -2.1); orient = _lib.bmMat3((C.c_float * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1))
    L.bmMarch(root.data_ptr(), C.byref(lo), C.byref(hi), rays.data_ptr(), buf.data_ptr(), 800, 200, 200, C.byref(eye), C.byref(orient), table.data_ptr())
    torch.cuda.synchronize()
    want = np.load(os.path.join(GOLDEN, "pixels_suzanne_200x200.npz"))["pixels"].reshape(-1)
    assert np.array_equal(buf.cpu().numpy().view(np.uint32), want)
    L.bmClear(buf.data_ptr(), 800, 200, 200, 7)
    torch.cuda.synchronize()
    assert (buf.cpu().numpy() == 7).all()
    # SceneHash.cpp's call sequence (TREE_TYPE == HASH): the scene is keyed by `cells` and renders the same pixels
    cells = torch.zeros(64, dtype=torch.uint8, device=dev)
    L.bmResetSpace(cells.data_ptr())
    L.bmInsertMeshInSpace(Pd.data_ptr(), Id.data_ptr(), Id.numel(), None, 0, cells.data_ptr())
    L.bmMarchSpace(rays.data_ptr(), buf.data_ptr(), 800, 200, 200, C.byref(eye), C.byref(orient), table.data_ptr(), cells.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy().view(np.uint32), want)
    L.bmMarchProgressive(None, 0, None, None, None, None, None, None, None, None)      # empty in the reference too (Trace2.cu:188-215)
    # an unknown root (nothing inserted) renders the miss colour
    root2 = torch.zeros(32, dtype=torch.uint8, device=dev)
    L.bmMarch(root2.data_ptr(), C.byref(lo), C.byref(hi), rays.data_ptr(), buf.data_ptr(), 800, 200, 200, C.byref(eye), C.byref(orient), table.data_ptr())
    torch.cuda.synchronize()
    assert (buf.cpu().numpy().view(np.uint32) == 255 << 8).all()


def test_hash_keys_on_device(golden, beam_lib, cuda_device):
    import torch
    z = np.load(os.path.join(GOLDEN, "hashkeys.npz"))
    pts = torch.from_numpy(z["pts"]).to(cuda_device)
    h = torch.empty(pts.shape[0], dtype=torch.int32, device=cuda_device); c = torch.empty((pts.shape[0], 3), dtype=torch.int32, device=cuda_device)
    assert beam_lib.bmHashKeys(pts.data_ptr(), pts.shape[0], h.data_ptr(), c.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(h.cpu().numpy().view(np.uint32), z["hash"]) and np.array_equal(c.cpu().numpy(), z["cells"])
    xs = np.array([[x, 0.0, 0.0] for x, _ in golden["hash"]["bmMap"]], np.float32)
    pts = torch.from_numpy(xs).to(cuda_device); c = torch.empty((xs.shape[0], 3), dtype=torch.int32, device=cuda_device)
    assert beam_lib.bmHashKeys(pts.data_ptr(), xs.shape[0], None, c.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert list(c.cpu().numpy()[:, 0]) == [w for _, w in golden["hash"]["bmMap"]]


def test_progressive_frame(oracle_lib, beam_lib, cuda_device):
    """ICamera::traceSceneProgressive (what SceneProgressive.cpp set out to be): N jittered samples per pixel, one diffuse bounce per
    hit, accumulated and resolved.  Checked three ways: every traced batch against the oracle on the SAME ray buffer, the accumulation
    against numpy, and the host-class frame against the same steps issued through the C ABI."""
    import ctypes as C
    import torch
    from oracle.bindings import OracleScene
    from raytracercuda_b200 import host as H, synth, workloads
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    wl = workloads.get("c2small")
    W, Hh = 160, 96
    n = W * Hh
    SPP, SEED = 3, 11
    # host-class path
    scene = H.IScene.create(); m = H.IMesh.create()
    assert m.setIndices(wl.indices, wl.indices.size) == 0 and m.setVertexData(wl.positions, wl.positions.shape[0], 3, 0) == 0
    assert m.setVertexData(wl.normals, wl.normals.shape[0], 3, 1) == 0
    scene.addMesh(m); scene.updateGPUScene()
    cam = H.ICamera.create(); assert cam.setInitialRays(W, Hh, -1, 1, -1, 1, 1) == 0
    rt = H.IRenderTarget.createHeadless(W, Hh); assert rt.lock() == 0
    assert cam.traceSceneProgressive(synth.EYE, synth.IDENTITY, scene, 0)[0] == H.ERROR_INVALID_PARAMETER
    err, rays_traced = cam.traceSceneProgressive(synth.EYE, synth.IDENTITY, scene, SPP, SEED)
    assert err == 0
    frame = rt.readPixels().reshape(-1)
    # the same steps through the C ABI, with the oracle looking at every batch
    t = Tree(cuda_device); t.add_mesh(wl.positions, wl.indices); t.build()
    s = OracleScene(oracle_lib); s.add_mesh(wl.positions, wl.indices, wl.normals); s.build()
    table = torch.from_numpy(cam.initialRays()).to(cuda_device)
    acc = torch.zeros(n, dtype=torch.float32, device=cuda_device); cov = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    acc_np = np.zeros(n, np.float32); cov_np = np.zeros(n, bool)
    jx, jy = 2.0 / W, 2.0 / Hh
    total = 0
    prev = None
    for sidx in range(SPP):
        r1 = t.camera_rays(table, synth.EYE, synth.IDENTITY, width=W, jitter=(jx, jy), sample=sidx, seed=SEED)
        R1 = r1.cpu().numpy()
        if prev is not None:
            assert np.abs(R1[:, 3:] - prev[:, 3:]).max() > 1e-5                      # samples differ ...
        prev = R1
        centre = cam.initialRays()
        ang = np.arccos(np.clip((R1[:, 3:] * centre).sum(1), -1, 1))
        assert ang.max() < 1.5 * np.hypot(jx, jy) and np.abs(np.linalg.norm(R1[:, 3:], axis=1) - 1).max() < 1e-5     # ... but stay inside their pixel
        h1 = t.trace_rays(r1, 0)
        a = hits_to_numpy(h1); o1 = s.trace_rays(R1[:, :3], R1[:, 3:], 0, 4)
        ok = o1["tie"] == 0
        assert (a["tri"][ok] == o1["tri"][ok]).mean() >= 0.9999
        r2, src, cnt = t.bounce_rays(r1, h1, sample=sidx, bounce=1, seed=SEED)
        nb = int(cnt.cpu()[0])
        h2 = torch.empty((n, 4), dtype=torch.float32, device=cuda_device)
        # bounce rays are traced as CLOSEST hits with t >= 0: in the reference's any-sign semantics a bounce ray "hits" the plane of its own
        # source triangle at t = -1e-4 / cos in the first leaf it enters, and every pixel would come out as flat ambient grey
        assert beam_lib.bmTraceRaysIndirect(t.h, r2.data_ptr(), cnt.data_ptr(), n, 1, h2.data_ptr(), None) == 0
        b = hits_to_numpy(h2[:nb]); R2 = r2[:nb].cpu().numpy(); o2 = s.trace_rays(R2[:, :3], R2[:, 3:], 1, 4)
        ok2 = o2["tie"] == 0
        assert (b["tri"][ok2] == o2["tri"][ok2]).mean() >= 0.9999
        both2 = ok2 & (o2["tri"] != 0xFFFFFFFF) & (b["tri"] == o2["tri"])
        assert both2.sum() == 0 or (np.abs(b["t"][both2] - o2["t"][both2]) <= 1e-5 * np.abs(o2["t"][both2])).all()
        assert (b["t"][b["tri"] != 0xFFFFFFFF] >= 0).all()
        # the scene is a displaced sphere seen from outside: most cosine-weighted bounce rays must escape (ambient occlusion, not self-hits)
        assert (b["tri"] == 0xFFFFFFFF).mean() > 0.5
        assert beam_lib.bmProgressiveAccumulate(h1.data_ptr(), n, h2.data_ptr(), src.data_ptr(), cnt.data_ptr(), 0.25, acc.data_ptr(), cov.data_ptr(), None) == 0
        hit = a["tri"] != 0xFFFFFFFF
        acc_np[hit] += np.float32(0.25); cov_np |= hit
        esc = src.cpu().numpy()[:nb][b["tri"] == 0xFFFFFFFF]
        acc_np[esc] += np.float32(0.75)
        total += n + nb
    px = torch.empty(n, dtype=torch.int32, device=cuda_device)
    assert beam_lib.bmProgressiveResolve(acc.data_ptr(), cov.data_ptr(), n, SPP, px.data_ptr(), None) == 0
    torch.cuda.synchronize()
    got = px.cpu().numpy().view(np.uint32)
    g = (np.clip(acc_np * np.float32(1.0 / SPP), 0, 1) * np.float32(255)).astype(np.uint32)
    want = np.where(cov_np, (g << 16) | (g << 8) | g, np.uint32(255 << 8)).astype(np.uint32)
    assert np.array_equal(got, want)
    assert np.array_equal(frame, got) and rays_traced == total
    # ambient occlusion of a displaced sphere: most bounce rays escape (pixels well above the ambient-only 0.25), some are occluded
    assert cov_np.sum() > 1000 and (acc_np[cov_np] > 0.3 * SPP).mean() > 0.5 and (acc_np[cov_np] < SPP - 0.5).any()
    assert rt.unlock() == 0
    t.close(); s.close()


def test_scene_lifetime_and_late_mesh_changes(meshes, beam_lib, cuda_device):
    """A destroyed scene takes its tree with it (its 32-byte root allocation is readily recycled by cudaMalloc: the next scene must not find
    a stale tree), and buffers replaced on an already-added mesh reach the scene's pointer tables before the next frame."""
    from raytracercuda_b200 import host as H
    L = beam_lib
    W = Hh = 64
    cam = H.ICamera.create(); assert cam.setInitialRays(W, Hh, -1, 1, -1, 1, 1) == 0
    rt = H.IRenderTarget.createHeadless(W, Hh); assert rt.lock() == 0
    eye, ori = [0, 0, -2.1], np.eye(3, dtype=np.float32)
    for round_ in range(6):
        scene = _scene(meshes, ["suzanne"]); scene.updateGPUScene()
        assert cam.traceScene(eye, ori, scene) == 0
        first = rt.readPixels()
        assert (first != 255 << 8).sum() > 100
        root_tree = scene.tree_handle()
        assert root_tree
        del scene                                        # ~Scene -> bmReleaseScene: the tree and its buffers are gone
        fresh = H.IScene.create()                        # very likely gets the same root address back
        assert not fresh.tree_handle()
        assert cam.traceScene(eye, ori, fresh) == 0      # nothing inserted: every pixel misses, as in the reference
        assert (rt.readPixels() == 255 << 8).all()
        del fresh
    # replacing the normals of an added mesh: the shade must read the new buffer (and never the freed one)
    P, I, N = meshes["suzanne"]
    scene = H.IScene.create(); m = H.IMesh.create()
    assert m.setIndices(I, I.size) == 0 and m.setVertexData(P, P.shape[0], 3, 0) == 0 and m.setVertexData(N, N.shape[0], 3, 1) == 0
    scene.addMesh(m); scene.updateGPUScene()
    assert cam.traceScene(eye, ori, scene) == 0
    a = rt.readPixels()
    flat = np.tile(np.array([[0, 0, 1]], np.float32), (P.shape[0], 1))
    assert m.setVertexData(flat, P.shape[0], 3, 1) == 0
    assert cam.traceScene(eye, ori, scene) == 0
    b = rt.readPixels()
    hit = a != 255 << 8
    assert np.array_equal(hit, b != 255 << 8) and (b[hit] == 255 << 16).all() and (a[hit] != 255 << 16).any()
    assert L.bmStickyError() == 0
    assert rt.unlock() == 0


def test_part1_failures_are_sticky_not_fatal(beam_lib, cuda_device):
    """the reference's launchers return void and exit(1) on failure (CudaComon.cuh:59-68); this library keeps the first failure instead"""
    import torch
    from raytracercuda_b200._lib import bmVec3
    L = beam_lib
    L.bmClearStickyError()
    root = torch.zeros(8, dtype=torch.int32, device=cuda_device)
    pos = torch.zeros(9, dtype=torch.float32, device=cuda_device); idx = torch.arange(3, dtype=torch.int32, device=cuda_device)
    lo, hi = bmVec3(-30, -30, -30), bmVec3(30, 30, 30)
    L.bmInsertMeshInTree(pos.data_ptr(), idx.data_ptr(), 4, 0, C.byref(lo), C.byref(hi), root.data_ptr(), None, None, None, None)      # 4 indices: not triangles
    assert L.bmStickyError() == 2 and b"bmInsertMeshInTree" in L.bmStickyErrorString()
    bad_lo, bad_hi = bmVec3(0, 0, 0), bmVec3(1, 1e-9, 1)
    L.bmInsertMeshInTree(pos.data_ptr(), idx.data_ptr(), 3, 0, C.byref(bad_lo), C.byref(bad_hi), root.data_ptr(), None, None, None, None)
    assert L.bmStickyError() == 2                         # the FIRST failure is kept
    L.bmClearStickyError()
    assert L.bmStickyError() == 0
    L.bmReleaseScene(root.data_ptr())
    L.bmReleaseScene(root.data_ptr())                     # unknown roots are ignored
    assert L.bmTreeFromRoot(root.data_ptr()) is None


def test_pair_total_beyond_32_bits_is_a_capacity_error(beam_lib, cuda_device):
    """4,600 world-spanning triangles reach ~1 M leaf cells each: the pair total passes 2^32 and must come back as BM_ERROR_CAPACITY,
    not as wrapped offsets"""
    from raytracercuda_b200.tree import Tree
    n = 4600
    z = np.linspace(-20, 20, n).astype(np.float32)
    P = np.zeros((n, 3, 3), np.float32)
    P[:, 0] = [-29.9, -29.9, 0]; P[:, 1] = [29.9, -29.9, 0]; P[:, 2] = [0, 29.9, 0]
    P[:, :, 2] = z[:, None]
    t = Tree(cuda_device); t.add_mesh(P.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32))
    r = beam_lib.bmTreeBuild(t.h)
    assert r == 103 and b"pairs" in beam_lib.bmLastErrorString()
    t.close()
