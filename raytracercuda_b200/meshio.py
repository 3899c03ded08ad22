"""Wavefront OBJ reader that preserves what parity depends on.

The reference imports meshes with Assimp (TestProgram/Model.cpp:26-54: Triangulate | JoinIdenticalVertices |
CalcTangentSpace | SortByPType, one IMesh per aiMesh = per material).  Assimp is a Windows binary in the reference tree, so
this module defines the import for this repo:

  * vertex positions are kept exactly as parsed (float32), in file order;
  * faces keep file order; polygons are fan-triangulated (v0, v[i], v[i+1]);
  * `split_by_material=True` yields one mesh per `usemtl` group in order of first use (the per-material split Assimp performs);
    the default is one mesh for the whole file, which is the form the golden tree fingerprints are defined on;
  * normals, when the file has `vn` records, are de-indexed to per-position normals (first use wins) so that the
    reference's shading path (BuildTree.cu:486-496, VERTEX_DATA_NORMAL indexed by the position index) has data.
"""
from __future__ import annotations

import io
import zipfile
from dataclasses import dataclass

import numpy as np


@dataclass
class TriMesh:
    positions: np.ndarray          # (V,3) float32
    indices: np.ndarray            # (F,3) uint32
    normals: np.ndarray | None     # (V,3) float32 or None
    name: str = ""
    uvs: np.ndarray | None = None  # (V,2) float32 or None (native loader only)
    texture: str | None = None     # map_Kd path of the mesh's material (native loader only)

    @property
    def num_tris(self) -> int:
        return int(self.indices.shape[0])


def _open_text(path: str):
    if path.endswith(".zip"):
        z = zipfile.ZipFile(path)
        names = [n for n in z.namelist() if n.lower().endswith(".obj")]
        return io.TextIOWrapper(z.open(names[0]), encoding="latin-1")
    return open(path, "r", encoding="latin-1")


def load_obj(path: str, split_by_material: bool = False) -> list[TriMesh]:
    pos: list[tuple[float, float, float]] = []
    nrm: list[tuple[float, float, float]] = []
    groups: dict[str, list[tuple[int, int, int]]] = {}
    order: list[str] = []
    cur = "__default__"
    vn_of_v: dict[int, int] = {}
    with _open_text(path) as f:
        for line in f:
            if not line or line[0] == "#":
                continue
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "v":
                pos.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "vn":
                nrm.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "usemtl":
                cur = t[1] if split_by_material and len(t) > 1 else "__default__"
            elif k == "f":
                vs = []
                for c in t[1:]:
                    p = c.split("/")
                    vi = int(p[0])
                    vi = vi - 1 if vi > 0 else len(pos) + vi
                    vs.append(vi)
                    if len(p) >= 3 and p[2]:
                        ni = int(p[2])
                        ni = ni - 1 if ni > 0 else len(nrm) + ni
                        vn_of_v.setdefault(vi, ni)
                if cur not in groups:
                    groups[cur] = []
                    order.append(cur)
                g = groups[cur]
                for i in range(1, len(vs) - 1):
                    g.append((vs[0], vs[i], vs[i + 1]))
    P = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    N = None
    if nrm and vn_of_v:
        NA = np.asarray(nrm, dtype=np.float32).reshape(-1, 3)
        N = np.zeros_like(P)
        N[:, 2] = 1.0
        vi = np.fromiter(vn_of_v.keys(), dtype=np.int64)
        ni = np.fromiter(vn_of_v.values(), dtype=np.int64)
        N[vi] = NA[ni]
    out = []
    for name in order:
        I = np.asarray(groups[name], dtype=np.uint32).reshape(-1, 3)
        out.append(TriMesh(P, I, N, name))
    return out


def compute_vertex_normals(positions: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """Area-weighted smooth normals (for meshes / soups that ship none; the reference shader needs slot 1)."""
    P = positions.astype(np.float64)
    I = indices.astype(np.int64)
    fn = np.cross(P[I[:, 1]] - P[I[:, 0]], P[I[:, 2]] - P[I[:, 0]])
    N = np.zeros_like(P)
    for k in range(3):
        np.add.at(N, I[:, k], fn)
    l = np.linalg.norm(N, axis=1, keepdims=True)
    N = np.where(l > 0, N / np.maximum(l, 1e-300), np.array([0.0, 0.0, 1.0]))
    return N.astype(np.float32)


def load_bmp_native(path: str) -> np.ndarray:
    """beamBmpLoad: (height, width) uint32 texels 0x00RRGGBB, row 0 = top"""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    w, h = C.c_uint32(0), C.c_uint32(0)
    p = L.beamBmpLoad(path.encode(), C.byref(w), C.byref(h))
    if not p:
        raise _lib.BeamError("cannot load BMP " + path)
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()
    finally:
        L.beamImageFree(p)


def load_obj_native(path: str, split_by_material: bool = False) -> list[TriMesh]:
    """the library's own C++ loader (csrc/objload.cpp, beamObjLoad) -- same rules as load_obj above"""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    h = L.beamObjLoad(path.encode(), int(split_by_material))
    try:
        err = L.beamObjError(h)
        if err:
            raise _lib.BeamError(err.decode())
        V = L.beamObjNumVertices(h)
        P = np.ctypeslib.as_array(L.beamObjPositions(h), shape=(V, 3)).copy()
        N = np.ctypeslib.as_array(L.beamObjNormals(h), shape=(V, 3)).copy() if L.beamObjHasNormals(h) else None
        UV = np.ctypeslib.as_array(L.beamObjUVs(h), shape=(V, 2)).copy() if L.beamObjHasUVs(h) else None
        out = []
        for m in range(L.beamObjNumMeshes(h)):
            n = L.beamObjMeshNumIndices(h, m)
            I = np.ctypeslib.as_array(L.beamObjMeshIndices(h, m), shape=(n,)).copy().reshape(-1, 3)
            tex = L.beamObjMeshTexture(h, m)
            out.append(TriMesh(P, I, N, L.beamObjMeshName(h, m).decode(), UV, tex.decode() if tex else None))
        return out
    finally:
        L.beamObjFree(h)
