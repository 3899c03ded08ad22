"""The UV / material path: `vt` + mtllib/map_Kd + BMP import on the CPU tier; textured deferred shading
(bmShadeHits, ICamera::traceSceneTextured) against the oracle's statement of the same rule on the GPU tier."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN

EYE = [0.0, 0.0, -2.1]
ID3 = np.eye(3, dtype=np.float32)


def _write_bmp24(path, rgb, top_down=False):
    """rgb: (h, w, 3) uint8, row 0 = top"""
    h, w, _ = rgb.shape
    stride = (w * 3 + 3) & ~3
    rows = rgb if top_down else rgb[::-1]
    data = b"".join(bytes(r[:, ::-1].reshape(-1)) + b"\0" * (stride - w * 3) for r in rows)
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", 54 + len(data), 0, 0, 54))
        f.write(struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, 24, 0, len(data), 2835, 2835, 0, 0))
        f.write(data)


def _quad_fixture(d):
    """a unit quad in front of the default camera, two material groups (left half textured, right half not)"""
    rgb = np.zeros((2, 3, 3), np.uint8)              # 3 wide: exercises the 4-byte row padding
    rgb[0, 0] = (255, 0, 0); rgb[0, 1] = (0, 255, 0); rgb[0, 2] = (0, 0, 255)
    rgb[1, 0] = (255, 255, 0); rgb[1, 1] = (0, 255, 255); rgb[1, 2] = (255, 255, 255)
    _write_bmp24(os.path.join(d, "tex.bmp"), rgb)
    _write_bmp24(os.path.join(d, "tex_td.bmp"), rgb, top_down=True)
    open(os.path.join(d, "quad.mtl"), "w").write("newmtl painted\nKd 1 1 1\nmap_Kd tex.bmp\n\nnewmtl plain\nKd 1 1 1\n")
    open(os.path.join(d, "quad.obj"), "w").write(
        "mtllib quad.mtl\n"
        "v -1 -1 1\nv 0 -1 1\nv 0 1 1\nv -1 1 1\nv 1 -1 1\nv 1 1 1\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 2 0\nvt 2 1\n"
        "vn 0 0 -1\n"
        "usemtl painted\nf 1/1/1 2/2/1 3/3/1 4/4/1\n"
        "usemtl plain\nf 2/2/1 5/5/1 6/6/1 3/3/1\n")
    return rgb


def test_obj_uv_mtl_bmp_import(tmp_path, beam_lib):
    from raytracercuda_b200 import meshio
    rgb = _quad_fixture(str(tmp_path))
    ms = meshio.load_obj_native(str(tmp_path / "quad.obj"), split_by_material=True)
    assert [m.name for m in ms] == ["painted", "plain"] and [m.num_tris for m in ms] == [2, 2]
    assert np.array_equal(ms[0].uvs, np.array([[0, 0], [1, 0], [1, 1], [0, 1], [2, 0], [2, 1]], np.float32))
    assert ms[0].texture == str(tmp_path / "tex.bmp") and ms[1].texture is None
    want = (rgb[..., 0].astype(np.uint32) << 16) | (rgb[..., 1].astype(np.uint32) << 8) | rgb[..., 2]
    assert np.array_equal(meshio.load_bmp_native(ms[0].texture), want)                      # bottom-up rows, padded stride
    assert np.array_equal(meshio.load_bmp_native(str(tmp_path / "tex_td.bmp")), want)       # top-down rows
    with pytest.raises(Exception):
        meshio.load_bmp_native(str(tmp_path / "quad.obj"))
    # without the split the file is one mesh and carries no material
    one = meshio.load_obj_native(str(tmp_path / "quad.obj"), split_by_material=False)
    assert len(one) == 1 and one[0].num_tris == 4 and one[0].uvs is not None


def test_oracle_material_rule(oracle_lib):
    """the rule on a case worked by hand: uv (0.5, 0.25) of a 2x2 texture, row 0 = top, v up -> bottom-right texel; n = (0,0,-1) -> s = 1"""
    from oracle.bindings import OracleScene
    s = OracleScene(oracle_lib)
    P = np.array([[-1, -1, 1], [1, -1, 1], [0, 1, 1]], np.float32)
    s.add_mesh(P, np.array([[0, 1, 2]], np.uint32), np.tile(np.array([[0, 0, -0.5]], np.float32), (3, 1)))
    miss = np.array([0xFFFFFFFF], np.uint32)
    assert s.shade_hits([0.5], [0.25], [0], miss)[0] == 255 << 8
    plain = s.shade_hits([0.5], [0.25], [0], [0])[0]
    assert plain == 255 << 16                                                    # the reference's pixel without a texture
    s.set_mesh_uv(0, np.array([[0, 0], [1, 0], [0, 1]], np.float32))            # uv = (u, v) of the hit
    s.set_mesh_texture(0, np.array([[0x102030, 0x405060], [0x708090, 0xA0B0C0]], np.uint32))
    assert s.shade_hits([0.5], [0.25], [0], [0])[0] == 0xA0B0C0                  # x = 1, y = (1 - 0.25) * 2 = 1
    assert s.shade_hits([0.25], [0.75], [0], [0])[0] == 0x102030                 # x = 0, y = 0
    assert s.shade_hits([0.0], [0.0], [0], [0])[0] == 0x708090                   # tv = 1 - 0 = 1 -> clamped to the last row
    s.close()


@pytest.mark.gpu
def test_shade_hits_matches_oracle_on_f16_materials(meshes, beam_lib, oracle_lib, cuda_device):
    """f16 as the two meshes of its material groups with the reference's own UVs and (decimated) map_Kd images: bmShadeHits pixels are
    bit-equal to the oracle's; mesh 1 is left without texture and must show the reference's pixel"""
    import torch
    from oracle.bindings import OracleScene
    from raytracercuda_b200 import _lib
    from raytracercuda_b200.host import ICamera
    from raytracercuda_b200.tree import Tree, hits_to_numpy, make_mesh_table
    L = beam_lib
    dev = cuda_device
    z = np.load(os.path.join(GOLDEN, "f16_materials.npz"))
    P, _, N = meshes["f16"]
    uv = z["uv"]
    groups = [z["idx_F16s"], z["idx_F16t"]]
    assert uv.shape[0] == P.shape[0]
    Pd = torch.from_numpy(P).to(dev); Nd = torch.from_numpy(N).to(dev); UVd = torch.from_numpy(uv).to(dev)
    Ids = [torch.from_numpy(g.view(np.int32)).to(dev).reshape(-1) for g in groups]
    tree = Tree(dev)
    for I in Ids:
        tree.add_mesh(Pd, I)
    tree.build()
    table = make_mesh_table([(Pd, Ids[0], Nd, UVd), (Pd, Ids[1], Nd, UVd)], dev)
    W = H = 300
    cam = ICamera.create(); cam.setInitialRays(W, H, -1, 1, -1, 1, 1)
    rays = torch.from_numpy(cam.initialRays()).to(dev)
    hits = torch.empty((W * H, 4), dtype=torch.float32, device=dev); pix = torch.zeros(W * H, dtype=torch.int32, device=dev)
    f3 = (C.c_float * 3)(*EYE); f9 = (C.c_float * 9)(*ID3.reshape(-1))
    _lib.check(L.bmTraceCamera(tree.h, rays.data_ptr(), W * H, W, f3, f9, 0, hits.data_ptr(), pix.data_ptr(), table.data_ptr(), None), "bmTraceCamera")
    ref_pixels = pix.cpu().numpy().view(np.uint32).copy()
    tex0 = torch.from_numpy(z["tex_F16s"].view(np.int32)).to(dev)
    texs = (_lib.bmTexture * 1)(); texs[0].texels = tex0.data_ptr(); texs[0].width = tex0.shape[1]; texs[0].height = tex0.shape[0]
    mats = (_lib.bmMaterial * 2)(); mats[0].texture = 0; mats[1].texture = -1
    d_tex = torch.from_numpy(np.frombuffer(bytes(texs), np.uint8).copy()).to(dev); d_mat = torch.from_numpy(np.frombuffer(bytes(mats), np.uint8).copy()).to(dev)
    _lib.check(L.bmShadeHits(tree.h, hits.data_ptr(), W * H, table.data_ptr(), d_mat.data_ptr(), d_tex.data_ptr(), 1, pix.data_ptr()), "bmShadeHits")
    torch.cuda.synchronize()
    got = pix.cpu().numpy().view(np.uint32)
    h = hits_to_numpy(hits)
    bases = np.array([0, groups[0].shape[0]], np.uint32)
    hit = h["tri"] != 0xFFFFFFFF
    mesh = np.where(hit, (h["tri"] >= bases[1]).astype(np.uint32), 0).astype(np.uint32)
    face = np.where(hit, h["tri"] - bases[mesh], 0xFFFFFFFF).astype(np.uint32)
    o = OracleScene(oracle_lib)
    for g in groups:
        o.add_mesh(P, g, N)
    o.set_mesh_uv(0, uv); o.set_mesh_uv(1, uv)
    o.set_mesh_texture(0, z["tex_F16s"])
    want = o.shade_hits(h["u"], h["v"], mesh, face)
    o.close()
    assert hit.sum() > 2000 and (mesh[hit] == 1).sum() > 50
    assert np.array_equal(got, want)
    # the untextured mesh and the misses keep the reference's pixels; the textured mesh shows colour
    keep = ~hit | (mesh == 1)
    assert np.array_equal(got[keep], ref_pixels[keep])
    assert ((got[hit & (mesh == 0)] & 0xFFFF) != 0).any()
    # no materials at all: exactly bmTraceCamera's pixels
    _lib.check(L.bmShadeHits(tree.h, hits.data_ptr(), W * H, table.data_ptr(), None, None, 0, pix.data_ptr()), "bmShadeHits")
    torch.cuda.synchronize()
    assert np.array_equal(pix.cpu().numpy().view(np.uint32), ref_pixels)


@pytest.mark.gpu
def test_host_api_textured_scene_from_obj(tmp_path, beam_lib, oracle_lib, cuda_device):
    """beamSceneLoadObj (Model::load) -> meshes with UVs and the material's BMP -> ICamera::traceSceneTextured, against the oracle;
    ICamera::traceScene on the same scene still shows the reference's pixels"""
    from oracle.bindings import OracleScene
    from raytracercuda_b200 import meshio
    from raytracercuda_b200.host import ICamera, IRenderTarget, IScene
    _quad_fixture(str(tmp_path))
    L = beam_lib
    L.beamSceneLoadObj.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    scene = IScene.create()
    assert L.beamSceneLoadObj(scene.h, str(tmp_path / "quad.obj").encode(), 1) == 2
    scene.updateGPUScene()
    W = H = 96
    cam = ICamera.create(); assert cam.setInitialRays(W, H, -1, 1, -1, 1, 1) == 0
    rt = IRenderTarget.createHeadless(W, H); assert rt.lock() == 0
    assert cam.traceScene(EYE, ID3, scene) == 0
    plain = rt.readPixels().reshape(-1).copy()
    assert cam.traceSceneTextured(EYE, ID3, scene) == 0
    got = rt.readPixels().reshape(-1).copy()
    err, hits = cam.traceSceneHits(EYE, ID3, scene)
    assert err == 0
    ms = meshio.load_obj_native(str(tmp_path / "quad.obj"), split_by_material=True)
    o = OracleScene(oracle_lib)
    for k, m in enumerate(ms):
        o.add_mesh(m.positions, m.indices, m.normals)
        o.set_mesh_uv(k, m.uvs)
    o.set_mesh_texture(0, meshio.load_bmp_native(ms[0].texture))
    tri = hits["tri"]; hit = tri != 0xFFFFFFFF
    mesh = np.where(hit, (tri >= 2).astype(np.uint32), 0).astype(np.uint32)
    face = np.where(hit, tri - 2 * mesh, 0xFFFFFFFF).astype(np.uint32)
    want = o.shade_hits(hits["u"], hits["v"], mesh, face)
    o.close()
    assert hit.sum() > 800                      # the 2x2 quad at distance 3.1 fills about a tenth of the 96x96 frame
    assert np.array_equal(got, want)
    assert np.array_equal(got[~hit | (mesh == 1)], plain[~hit | (mesh == 1)])
    colours = set(int(x) for x in np.unique(got[hit & (mesh == 0)]))
    assert {0xFF0000, 0x00FF00, 0x0000FF, 0xFFFF00, 0x00FFFF, 0xFFFFFF} <= colours          # all six texels of the BMP appear, unshaded (|n.z| = 1)
    assert rt.unlock() == 0
