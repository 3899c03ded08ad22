"""CPU tier: the library's C++ OBJ loader (csrc/objload.cpp) -- the replacement for the reference's Assimp import
(TestProgram/Model.cpp) -- on generated files, and against the Python loader / golden meshes on the reference's bundled assets."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _write(tmp_path, text, name="m.obj"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_rules(tmp_path, beam_lib):
    from raytracercuda_b200.meshio import load_obj, load_obj_native
    src = """# comment
mtllib x.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0.5 0.5 1.25
vn 0 0 1
vn 0 1 0
usemtl A
f 1/1/1 2/2/1 3/3/1 4/4/1
usemtl B
f 1//2 2//2 5//2
f -5 -4 -1
usemtl A
f 3 4 5
"""
    path = _write(tmp_path, src)
    for split in (False, True):
        a = load_obj(path, split_by_material=split); b = load_obj_native(path, split_by_material=split)
        assert len(a) == len(b) == (2 if split else 1)
        for x, y in zip(a, b):
            assert np.array_equal(x.positions, y.positions) and np.array_equal(x.indices, y.indices)
            assert np.array_equal(x.normals, y.normals)
    one = load_obj_native(path)[0]
    # quad -> fan (1,2,3),(1,3,4); negative indices resolve from the end; file order kept
    assert one.indices.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [0, 1, 4], [2, 3, 4]]
    assert one.normals[0].tolist() == [0, 0, 1] and one.normals[4].tolist() == [0, 1, 0]      # first use wins per position
    two = load_obj_native(path, split_by_material=True)
    assert [m.name for m in two] == ["A", "B"] and two[0].indices.tolist() == [[0, 1, 2], [0, 2, 3], [2, 3, 4]]
    assert beam_lib.beamObjError(beam_lib.beamObjLoad(b"/nonexistent.obj", 0)) is not None


@pytest.mark.parametrize("face", ["f -9//1 1//1 2//1", "f 2000000000//1 1//1 2//1", "f 0 1 2", "f 99999999999999999999 1 2", "f 1 2 -2000000000"])
def test_bad_face_indices_are_an_error_not_a_crash(tmp_path, beam_lib, face):
    """an index outside the tables read so far must come back as a loader error through the C ABI (no exception, no giant allocation)"""
    path = _write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\n" + face + "\n")
    h = beam_lib.beamObjLoad(path.encode(), 0)
    err = beam_lib.beamObjError(h)
    assert err is not None and b"out of range" in err
    assert beam_lib.beamObjNumVertices(h) == 0 and beam_lib.beamObjNumMeshes(h) == 0
    beam_lib.beamObjFree(h)


def test_dangling_normal_and_uv_references_are_ignored(tmp_path, beam_lib):
    from raytracercuda_b200.meshio import load_obj_native
    path = _write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0.5 0.25\nf 1/99/7 2/1/1 3/-5/-9\n")
    m = load_obj_native(path)[0]
    assert m.indices.tolist() == [[0, 1, 2]] and m.normals[1].tolist() == [0, 0, 1]


def test_ppm(tmp_path, beam_lib):
    px = np.array([[255 << 16, 255 << 8], [255, 0x123456]], np.uint32)
    out = str(tmp_path / "o.ppm")
    assert beam_lib.beamWritePPM(out.encode(), px.ctypes.data, 2, 2) == 0
    raw = open(out, "rb").read()
    assert raw.startswith(b"P6\n2 2\n255\n") and raw[-12:] == bytes([255, 0, 0, 0, 255, 0, 0, 0, 255, 0x12, 0x34, 0x56])


@pytest.mark.skipif(not os.path.isdir("/root/reference/Content"), reason="reference assets only exist in the build container")
def test_bundled_assets_match_golden_meshes(beam_lib, meshes):
    from raytracercuda_b200.meshio import load_obj_native
    for name in ("f16", "suzanne"):
        m = load_obj_native("/root/reference/Content/%s.obj" % name)[0]
        P, I, N = meshes[name]
        assert np.array_equal(m.positions, P) and np.array_equal(m.indices, I) and np.array_equal(m.normals, N)
    two = load_obj_native("/root/reference/Content/f16.obj", split_by_material=True)
    assert [int(t.indices.shape[0]) for t in two] == [3704, 352]          # aa.xml:56,58: grids of 15 and 2 blocks x 256
