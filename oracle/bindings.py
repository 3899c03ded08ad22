"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/_ref/libbeam_ref*.so (the unmodified reference in CUDA-0 mode)
and oracle/_build/liboracle.so (this repo's CPU restatement).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package never does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libbeam_ref.so"))


class RefLib:
    """The unmodified reference through its public Beam.h interface (oracle/ref_driver.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libbeam_ref.so"))
        L = self.lib
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_add_mesh.argtypes = [C.c_void_p, _f32p, C.c_void_p, C.c_uint32, _u32p, C.c_uint32]
        L.ref_scene_build.argtypes = [C.c_void_p]
        L.ref_scene_build.restype = C.c_double
        L.ref_scene_stats.argtypes = [C.c_void_p, _u64p]
        L.ref_scene_pairs.argtypes = [C.c_void_p, _u32p, _u32p, _u32p]
        L.ref_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_float] * 5 + [_f32p, _f32p, _u32p, C.POINTER(C.c_double)]
        L.ref_initial_rays.argtypes = [C.c_uint32, C.c_uint32] + [C.c_float] * 5 + [_f32p]
        L.ref_tri_intersect.argtypes = [_f32p] * 5 + [C.POINTER(C.c_float)] * 2
        L.ref_tri_intersect.restype = C.c_float
        L.ref_box_ray.argtypes = [_f32p] * 4
        L.ref_box_ray.restype = C.c_float
        L.ref_box_ray_nozero.argtypes = [_f32p] * 4
        L.ref_box_ray_nozero.restype = C.c_float
        L.ref_tri_box_overlap.argtypes = [_f32p] * 3
        L.ref_struct_sizes.argtypes = [_u32p]


class RefScene:
    def __init__(self, reflib: RefLib | None = None):
        self.R = reflib or RefLib()
        self.h = self.R.lib.ref_scene_create()
        self._keep = []

    def add_mesh(self, positions, indices, normals=None):
        P = np.ascontiguousarray(positions, dtype=np.float32)
        I = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        N = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        self._keep += [P, I, N]
        err = self.R.lib.ref_scene_add_mesh(self.h, P, None if N is None else N.ctypes.data, P.shape[0], I, I.size)
        assert err == 0, err

    def build(self) -> float:
        return self.R.lib.ref_scene_build(self.h)

    def stats(self) -> dict:
        s = np.zeros(8, dtype=np.uint64)
        self.R.lib.ref_scene_stats(self.h, s)
        k = ["pairs", "nodes", "leaves", "max_depth", "min_leaf_depth", "overflow_leaves", "max_leaf_faces"]
        return {n: int(v) for n, v in zip(k, s)}

    def pairs(self):
        n = self.stats()["pairs"]
        c = np.zeros(n, np.uint32); m = np.zeros(n, np.uint32); t = np.zeros(n, np.uint32)
        self.R.lib.ref_scene_pairs(self.h, c, m, t)
        return c, m, t

    def trace(self, W, H, eye, orient, frustum=(-1, 1, -1, 1, 1)):
        px = np.zeros(W * H, np.uint32)
        sec = C.c_double(0)
        e = np.ascontiguousarray(eye, np.float32); o = np.ascontiguousarray(orient, np.float32).reshape(-1)
        err = self.R.lib.ref_trace(self.h, W, H, *[float(x) for x in frustum], e, o, px, C.byref(sec))
        return err, px.reshape(H, W), sec.value

    def close(self):
        if self.h:
            self.R.lib.ref_scene_destroy(self.h)
            self.h = None


class RefHashLib:
    def __init__(self):
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libbeam_ref_hash.so"))
        L = self.lib
        L.ref_hash.argtypes = [C.c_uint32]; L.ref_hash.restype = C.c_uint32
        L.ref_hash3.argtypes = [C.c_uint32] * 3; L.ref_hash3.restype = C.c_uint32
        L.ref_hash3f.argtypes = [C.c_float] * 3; L.ref_hash3f.restype = C.c_uint32
        L.ref_map.argtypes = [C.c_float]; L.ref_map.restype = C.c_int32
        L.ref_cell_size.restype = C.c_uint32
        L.ref_max_hash_elements.restype = C.c_uint32


def fnv1a64_pairs(code: np.ndarray, tri: np.ndarray) -> int:
    """SURVEY.md section 8(c) tree fingerprint: FNV-1a-64 over the sorted (code31, triIdx) list, each value
    xor-then-multiplied as one 32-bit unit."""
    h = 1469598103934665603
    prime = 1099511628211
    mask = (1 << 64) - 1
    for c, t in zip(code.tolist(), tri.tolist()):
        h = ((h ^ c) * prime) & mask
        h = ((h ^ t) * prime) & mask
    return h


def oracle_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_build", "liboracle.so"))


class OracleLib:
    """oracle/oracle.cpp: this repo's CPU restatement of the reference build + traversal."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(HERE, "_build", "liboracle.so"))
        L = self.lib
        vp = C.c_void_p
        L.orc_tri_box_overlap.argtypes = [_f32p] * 3
        L.orc_tri_intersect.argtypes = [_f32p] * 5 + [C.POINTER(C.c_float)] * 2
        L.orc_tri_intersect.restype = C.c_float
        L.orc_box_ray.argtypes = [_f32p] * 4; L.orc_box_ray.restype = C.c_float
        L.orc_box_ray_nozero.argtypes = [_f32p] * 4; L.orc_box_ray_nozero.restype = C.c_float
        L.orc_hash.argtypes = [C.c_uint32]; L.orc_hash.restype = C.c_uint32
        L.orc_hash3.argtypes = [C.c_uint32] * 3; L.orc_hash3.restype = C.c_uint32
        L.orc_map.argtypes = [C.c_float]; L.orc_map.restype = C.c_int32
        L.orc_initial_rays.argtypes = [C.c_uint32, C.c_uint32] + [C.c_float] * 5 + [_f32p]
        L.orc_scene_create.argtypes = [vp, vp, C.c_float, C.c_uint32, C.c_uint32]; L.orc_scene_create.restype = vp
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_scene_add_mesh.argtypes = [vp, _f32p, vp, C.c_uint32, _u32p, C.c_uint32]
        L.orc_scene_build.argtypes = [vp]; L.orc_scene_build.restype = C.c_double
        L.orc_scene_build_mt.argtypes = [vp, C.c_int]; L.orc_scene_build_mt.restype = C.c_double
        L.orc_scene_stats.argtypes = [vp, _u64p]
        L.orc_scene_pairs.argtypes = [vp, _u32p, _u32p, _u32p]
        L.orc_scene_branches.argtypes = [vp, _u32p, _u32p]
        L.orc_trace_rays.argtypes = [vp, _f32p, _f32p, C.c_uint64, C.c_int] + [vp] * 7 + [vp, C.c_int]
        L.orc_trace_rays.restype = C.c_double
        L.orc_trace_camera.argtypes = [vp, _f32p, C.c_uint64, _f32p, _f32p, C.c_int] + [vp] * 7 + [vp, C.c_int]
        L.orc_trace_camera.restype = C.c_double

    def initial_rays(self, W, H, frustum=(-1, 1, -1, 1, 1)):
        out = np.zeros((H * W, 3), np.float32)
        err = self.lib.orc_initial_rays(W, H, *[float(x) for x in frustum], out)
        assert err == 0
        return out


class OracleScene:
    def __init__(self, lib: OracleLib | None = None, wmin=None, wmax=None, min_leaf=0.0, max_depth=0, max_faces=0):
        self.O = lib or OracleLib()
        a = None if wmin is None else np.ascontiguousarray(wmin, np.float32)
        b = None if wmax is None else np.ascontiguousarray(wmax, np.float32)
        self.h = self.O.lib.orc_scene_create(None if a is None else a.ctypes.data, None if b is None else b.ctypes.data,
                                             float(min_leaf), int(max_depth), int(max_faces))

    def add_mesh(self, positions, indices, normals=None):
        P = np.ascontiguousarray(positions, dtype=np.float32)
        I = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        N = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        err = self.O.lib.orc_scene_add_mesh(self.h, P, None if N is None else N.ctypes.data, P.shape[0], I, I.size)
        assert err == 0, err

    def build(self, nthreads: int = 1) -> float:
        """nthreads == 1: the serial insertion exactly as the reference's host mode performs it (the pinned path);
        nthreads > 1: the same tree from per-thread descents (orc_scene_build_mt; asserted equal in tests/test_oracle_golden.py)"""
        if nthreads > 1:
            return self.O.lib.orc_scene_build_mt(self.h, int(nthreads))
        return self.O.lib.orc_scene_build(self.h)

    def stats(self) -> dict:
        s = np.zeros(8, dtype=np.uint64)
        self.O.lib.orc_scene_stats(self.h, s)
        k = ["pairs", "nodes", "leaves", "max_depth", "min_leaf_depth", "overflow_leaves", "max_leaf_faces", "branches"]
        return {n: int(v) for n, v in zip(k, s)}

    def pairs(self):
        n = self.stats()["pairs"]
        c = np.zeros(n, np.uint32); m = np.zeros(n, np.uint32); t = np.zeros(n, np.uint32)
        self.O.lib.orc_scene_pairs(self.h, c, m, t)
        return c, m, t

    def branches(self):
        n = self.stats()["branches"]
        d = np.zeros(n, np.uint32); p = np.zeros(n, np.uint32)
        self.O.lib.orc_scene_branches(self.h, d, p)
        return d, p

    def _outs(self, n):
        return dict(t=np.zeros(n, np.float32), u=np.zeros(n, np.float32), v=np.zeros(n, np.float32), mesh=np.zeros(n, np.uint32),
                    tri=np.zeros(n, np.uint32), tie=np.zeros(n, np.uint32), pixels=np.zeros(n, np.uint32))

    def trace_rays(self, org, dirs, mode=0, nthreads=1):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3); dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        o = self._outs(n); cnt = np.zeros(2, np.uint64)
        sec = self.O.lib.orc_trace_rays(self.h, org, dirs, n, mode, *[o[k].ctypes.data for k in ("t", "u", "v", "mesh", "tri", "tie", "pixels")],
                                        cnt.ctypes.data, nthreads)
        o["seconds"] = sec; o["node_visits"] = int(cnt[0]); o["tri_tests"] = int(cnt[1])
        return o

    def trace_camera(self, initial_rays, eye, orient, mode=0, nthreads=1):
        R = np.ascontiguousarray(initial_rays, np.float32).reshape(-1, 3)
        n = R.shape[0]
        e = np.ascontiguousarray(eye, np.float32); m = np.ascontiguousarray(orient, np.float32).reshape(-1)
        o = self._outs(n); cnt = np.zeros(2, np.uint64)
        sec = self.O.lib.orc_trace_camera(self.h, R, n, e, m, mode, *[o[k].ctypes.data for k in ("t", "u", "v", "mesh", "tri", "tie", "pixels")],
                                          cnt.ctypes.data, nthreads)
        o["seconds"] = sec; o["node_visits"] = int(cnt[0]); o["tri_tests"] = int(cnt[1])
        return o

    def set_mesh_uv(self, mesh, uv):
        a = np.ascontiguousarray(uv, np.float32)
        self.O.lib.orc_scene_set_mesh_uv.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        assert self.O.lib.orc_scene_set_mesh_uv(self.h, int(mesh), a.ctypes.data, int(a.shape[1])) == 0

    def set_mesh_texture(self, mesh, texels):
        a = np.ascontiguousarray(texels, np.uint32)
        self.O.lib.orc_scene_set_mesh_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        assert self.O.lib.orc_scene_set_mesh_texture(self.h, int(mesh), a.ctypes.data, int(a.shape[1]), int(a.shape[0])) == 0

    def shade_hits(self, u, v, mesh, tri):
        """pixels of hit records under the product's material rule (bmShadeHits); tri = face index within its mesh, 0xFFFFFFFF = miss"""
        u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
        mesh = np.ascontiguousarray(mesh, np.uint32); tri = np.ascontiguousarray(tri, np.uint32)
        out = np.zeros(u.shape[0], np.uint32)
        self.O.lib.orc_shade_hits.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 5
        self.O.lib.orc_shade_hits.restype = None
        self.O.lib.orc_shade_hits(self.h, u.shape[0], u.ctypes.data, v.ctypes.data, mesh.ctypes.data, tri.ctypes.data, out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.O.lib.orc_scene_destroy(self.h)
            self.h = None


class RefGpuScene:
    """The reference's UNMODIFIED host classes in CUDA-1 mode (oracle/ref_gpu_driver.cpp, oracle/build_ref_cuda.sh).
    which = "refhost_b200": on top of this repo's libbeam_b200.so (drop-in proof);  "refcuda": on top of the reference's own
    kernels rebuilt for sm_100a (second baseline).  Needs a GPU."""

    def __init__(self, which: str):
        path = os.path.join(HERE, "_ref", "libbeam_%s.so" % which)
        self.lib = C.CDLL(path, mode=os.RTLD_NOW | os.RTLD_LOCAL | getattr(os, "RTLD_DEEPBIND", 0))
        L = self.lib
        vp = C.c_void_p
        L.refgpu_scene_create.restype = vp
        L.refgpu_scene_destroy.argtypes = [vp]
        L.refgpu_scene_add_mesh.argtypes = [vp, _f32p, vp, C.c_uint32, _u32p, C.c_uint32]
        L.refgpu_scene_build.argtypes = [vp]; L.refgpu_scene_build.restype = C.c_double
        L.refgpu_camera.argtypes = [vp, C.c_uint32, C.c_uint32] + [C.c_float] * 5
        L.refgpu_trace.argtypes = [vp, _f32p, _f32p, C.c_int, C.POINTER(C.c_int)]; L.refgpu_trace.restype = C.c_double
        L.refgpu_read_pixels.argtypes = [vp, _u32p]
        self.h = L.refgpu_scene_create()
        self._keep = []
        self.W = self.H = 0

    @staticmethod
    def available(which: str) -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libbeam_%s.so" % which))

    def add_mesh(self, positions, indices, normals=None):
        P = np.ascontiguousarray(positions, dtype=np.float32)
        I = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        N = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        self._keep += [P, I, N]
        err = self.lib.refgpu_scene_add_mesh(self.h, P, None if N is None else N.ctypes.data, P.shape[0], I, I.size)
        assert err == 0, err

    def build(self) -> float:
        return self.lib.refgpu_scene_build(self.h)

    def camera(self, W, H, frustum=(-1, 1, -1, 1, 1)):
        err = self.lib.refgpu_camera(self.h, W, H, *[float(x) for x in frustum])
        assert err == 0, err
        self.W, self.H = W, H

    def trace(self, eye, orient, repeat=1):
        e = np.ascontiguousarray(eye, np.float32); o = np.ascontiguousarray(orient, np.float32).reshape(-1)
        err = C.c_int(0)
        sec = self.lib.refgpu_trace(self.h, e, o, int(repeat), C.byref(err))
        assert err.value == 0, err.value
        return sec

    def pixels(self):
        px = np.zeros(self.W * self.H, np.uint32)
        assert self.lib.refgpu_read_pixels(self.h, px) == 0
        return px.reshape(self.H, self.W)

    def close(self):
        if self.h:
            self.lib.refgpu_scene_destroy(self.h)
            self.h = None
