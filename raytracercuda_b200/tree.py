"""Python view of the handle-based C ABI (include/beam_b200.h, Part 2).  torch provides device memory and streams only;
every computation happens in libbeam_b200.so's CUDA kernels."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BM_TRACE_CLOSEST, BM_TRACE_REFERENCE, check  # noqa: F401

MISS = 0xFFFFFFFF


def _dev_tensor(a, dtype, device):
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=dtype).contiguous()
    else:
        arr = np.ascontiguousarray(a)
        if arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        t = torch.from_numpy(arr).to(device=device, dtype=dtype).contiguous()
    return t


def _u32(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)


class Tree:
    """bmTree*: staged meshes -> build -> trace.  Mirrors SceneTree's use of bmResetScene / bmInsertMeshInTree / bmMarch
    (SceneTree.cpp:70-127) on the explicit entry points."""

    def __init__(self, device="cuda:0", world_min=None, world_max=None, min_leaf=0.0, max_depth=0):
        self.L = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.BeamError("beam_b200 needs a CUDA device: there is no CPU path")
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        torch.zeros(1, device=self.device)          # make sure the primary context exists before the library's first call
        h = C.c_void_p()
        wmin = None if world_min is None else np.ascontiguousarray(world_min, np.float32)
        wmax = None if world_max is None else np.ascontiguousarray(world_max, np.float32)
        check(self.L.bmTreeCreate(C.byref(h), None if wmin is None else wmin.ctypes.data, None if wmax is None else wmax.ctypes.data,
                                  float(min_leaf), int(max_depth)), "bmTreeCreate")
        self.h = h
        self._meshes = []        # keeps device tensors alive: the library stages pointers, like the reference
        self.use_stream(torch.cuda.current_stream(self.device))

    def use_stream(self, stream: torch.cuda.Stream):
        self.stream = stream
        check(self.L.bmTreeSetStream(self.h, C.c_void_p(stream.cuda_stream)), "bmTreeSetStream")

    def reset(self):
        check(self.L.bmTreeReset(self.h), "bmTreeReset")
        self._meshes = []

    def add_mesh(self, positions, indices, mesh_idx=None):
        P = _dev_tensor(positions, torch.float32, self.device).reshape(-1, 3)
        I = _dev_tensor(indices, torch.int32, self.device).reshape(-1)
        if mesh_idx is None:
            mesh_idx = len(self._meshes)
        check(self.L.bmTreeAddMesh(self.h, P.data_ptr(), I.data_ptr(), I.numel(), int(mesh_idx)), "bmTreeAddMesh")
        self._meshes.append((P, I))
        return mesh_idx

    def build(self):
        check(self.L.bmTreeBuild(self.h), "bmTreeBuild")

    def stats(self) -> dict:
        s = (C.c_uint64 * 8)()
        check(self.L.bmTreeStats(self.h, s), "bmTreeStats")
        k = ["triangles", "pairs", "leaves", "nodes", "depth", "max_leaf_faces", "overflow_leaves", "meshes"]
        return {n: int(v) for n, v in zip(k, s)}

    def root(self) -> int:
        r = C.c_uint32(0)
        check(self.L.bmTreeRoot(self.h, C.byref(r)), "bmTreeRoot")
        return int(r.value)

    def build_stages(self) -> dict:
        """stage times (microseconds) of the last build made after enable_stage_timing(True)"""
        w = self.L.bmTreeLastBuildStages(self.h).decode().split()
        return {w[i]: float(w[i + 1]) for i in range(0, len(w) - 1, 2)}

    def enable_stage_timing(self, on=True):
        check(self.L.bmTreeEnableStageTiming(self.h, int(bool(on))), "bmTreeEnableStageTiming")

    def set_bricks(self, mode: int = -1, bits: int = 0):
        """how trace_rays walks big trees (bmTreeSetBricks): mode -1 automatic, 0 direct traversal, 1 brick by brick; bits = depth of the brick grid"""
        _lib.check(self.L.bmTreeSetBricks(self.h, mode, bits), "bmTreeSetBricks")

    def node_layout(self) -> int:
        """BM_NODES_BINARY (128-byte binary records), BM_NODES_WIDE (64-byte 4-wide records) or BM_NODES_BOTH"""
        v = C.c_uint32(0)
        check(self.L.bmTreeNodeLayout(self.h, C.byref(v)), "bmTreeNodeLayout")
        return int(v.value)

    def world(self):
        """(world_min[3], world_max[3], min_leaf, max_depth) of the tree's grid"""
        a = np.zeros(3, np.float32); b = np.zeros(3, np.float32); ml = C.c_float(0); md = C.c_uint32(0)
        check(self.L.bmTreeWorld(self.h, a.ctypes.data, b.ctypes.data, C.byref(ml), C.byref(md)), "bmTreeWorld")
        return a, b, float(ml.value), int(md.value)

    def wide_slots(self):
        """decoded culling boxes (N,4,6) and slot references (N,4) of the wide records"""
        N = self.stats()["nodes"]
        b = torch.empty((N, 4, 6), dtype=torch.float32, device=self.device); r = torch.empty((N, 4), dtype=torch.int32, device=self.device)
        check(self.L.bmTreeCopyWideSlots(self.h, b.data_ptr(), r.data_ptr()), "bmTreeCopyWideSlots")
        self.stream.synchronize()
        return b.cpu().numpy(), _u32(r)

    def mesh_bases(self) -> np.ndarray:
        n = len(self._meshes)
        b = (C.c_uint32 * max(n, 1))()
        check(self.L.bmTreeMeshBases(self.h, b, n), "bmTreeMeshBases")
        return np.array(b[:n], dtype=np.uint32)

    def pairs(self):
        P = self.stats()["pairs"]
        code = torch.empty(P, dtype=torch.int32, device=self.device); tri = torch.empty(P, dtype=torch.int32, device=self.device)
        check(self.L.bmTreeCopyPairs(self.h, code.data_ptr(), tri.data_ptr()), "bmTreeCopyPairs")
        self.stream.synchronize()
        return _u32(code), _u32(tri)

    def topology(self):
        N = self.stats()["nodes"]
        t = [torch.empty(N, dtype=torch.int32, device=self.device) for _ in range(4)]
        check(self.L.bmTreeCopyTopology(self.h, *[x.data_ptr() for x in t]), "bmTreeCopyTopology")
        self.stream.synchronize()
        plen, prefix, c0, c1 = [_u32(x) for x in t]
        return plen, prefix, c0, c1

    def child_boxes(self) -> np.ndarray:
        N = self.stats()["nodes"]
        b = torch.empty((N, 12), dtype=torch.float32, device=self.device)
        check(self.L.bmTreeCopyChildBoxes(self.h, b.data_ptr()), "bmTreeCopyChildBoxes")
        self.stream.synchronize()
        return b.cpu().numpy()

    def trace_rays(self, rays, mode=BM_TRACE_REFERENCE, counters=False, hits_out=None, num=None):
        """rays: (n,6) float32 (origin, direction) = bmRay records. Returns the (n,4) hit tensor [t,u,v,tri-bits] on the device."""
        R = _dev_tensor(rays, torch.float32, self.device).reshape(-1, 6)
        n = R.shape[0] if num is None else int(num)
        hits = hits_out if hits_out is not None else torch.empty((n, 4), dtype=torch.float32, device=self.device)
        cnt = torch.zeros(4, dtype=torch.int64, device=self.device) if counters else None
        check(self.L.bmTraceRays(self.h, R.data_ptr(), n, int(mode), hits.data_ptr(), None if cnt is None else cnt.data_ptr()), "bmTraceRays")
        return (hits, cnt) if counters else hits

    def trace_camera(self, initial_rays, eye, orient, mode=BM_TRACE_REFERENCE, want_hits=True, pixels=None, mesh_table=None, counters=False, width=0):
        R = _dev_tensor(initial_rays, torch.float32, self.device).reshape(-1, 3)
        n = R.shape[0]
        e = (C.c_float * 3)(*[float(x) for x in eye])
        o = (C.c_float * 9)(*[float(x) for x in np.asarray(orient, np.float32).reshape(-1)])
        hits = torch.empty((n, 4), dtype=torch.float32, device=self.device) if want_hits else None
        cnt = torch.zeros(4, dtype=torch.int64, device=self.device) if counters else None
        check(self.L.bmTraceCamera(self.h, R.data_ptr(), n, int(width), e, o, int(mode), None if hits is None else hits.data_ptr(),
                                   None if pixels is None else pixels.data_ptr(), None if mesh_table is None else mesh_table.data_ptr(),
                                   None if cnt is None else cnt.data_ptr()), "bmTraceCamera")
        return hits, cnt

    def camera_rays(self, initial_rays_dev: torch.Tensor, eye, orient, width=0, jitter=(0.0, 0.0), sample=0xFFFFFFFF, seed=0, out=None):
        """bmCameraRays: (n,6) bmRay records for the camera table (device tensor (n,3))"""
        n = initial_rays_dev.shape[0]
        out = out if out is not None else torch.empty((n, 6), dtype=torch.float32, device=self.device)
        e = (C.c_float * 3)(*[float(x) for x in eye])
        o = (C.c_float * 9)(*[float(x) for x in np.asarray(orient, np.float32).reshape(-1)])
        check(self.L.bmCameraRays(initial_rays_dev.data_ptr(), n, int(width), e, o, float(jitter[0]), float(jitter[1]), int(sample), int(seed),
                                  out.data_ptr(), C.c_void_p(self.stream.cuda_stream)), "bmCameraRays")
        return out

    def bounce_rays(self, rays: torch.Tensor, hits: torch.Tensor, sample=0, bounce=1, seed=0, out=None, src=None, count=None):
        """bmBounceRays: compacted diffuse bounce rays for the hits; returns (rays_out, src_index, count tensor (device int32[1]))"""
        n = rays.shape[0]
        out = out if out is not None else torch.empty((n, 6), dtype=torch.float32, device=self.device)
        src = src if src is not None else torch.empty(n, dtype=torch.int32, device=self.device)
        count = count if count is not None else torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self.L.bmBounceRays(self.h, rays.data_ptr(), hits.data_ptr(), n, int(sample), int(bounce), int(seed), out.data_ptr(), src.data_ptr(),
                                  count.data_ptr()), "bmBounceRays")
        return out, src, count

    def view(self, stream: torch.cuda.Stream = None) -> "Tree":
        """bmTreeView: a second handle on this built tree (no copy) with its own stream and scratch, for concurrent launches"""
        out = Tree.__new__(Tree)
        out.L = self.L; out.device = self.device; out._meshes = []
        h = C.c_void_p()
        check(self.L.bmTreeView(self.h, C.byref(h)), "bmTreeView")
        out.h = h
        out._source = self                   # the arrays belong to the source tree
        out.use_stream(stream if stream is not None else torch.cuda.current_stream(self.device))
        return out

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.bmTreeDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hits_to_numpy(hits: torch.Tensor):
    """(n,4) device tensor -> dict of numpy arrays t,u,v (float32) and tri (uint32, MISS on a miss)."""
    h = hits.cpu().numpy()
    return {"t": h[:, 0].copy(), "u": h[:, 1].copy(), "v": h[:, 2].copy(), "tri": h[:, 3].copy().view(np.uint32)}


def make_mesh_table(meshes, device) -> torch.Tensor:
    """Device array of StaticMeshData (SharedTypes.h:10-18, 136 bytes each) from [(positions, indices, normals|None[, uvs|None])] device tensors."""
    n = len(meshes)
    arr = (_lib.bmStaticMeshData * n)()
    for i, m in enumerate(meshes):
        P, I, N = m[0], m[1], m[2]
        if len(m) > 3 and m[3] is not None:
            arr[i].m_vertexData[2] = m[3].data_ptr(); arr[i].m_vertexDataSizes[2] = m[3].shape[1]      # VERTEX_DATA_UV1
        arr[i].m_vertexData[0] = P.data_ptr(); arr[i].m_vertexDataSizes[0] = 3
        if N is not None:
            arr[i].m_vertexData[1] = N.data_ptr(); arr[i].m_vertexDataSizes[1] = 3
        arr[i].m_indices = I.data_ptr(); arr[i].m_numVertices = P.shape[0]; arr[i].m_numIndices = I.numel()
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    table = torch.from_numpy(raw).to(device)
    table._beam_keepalive = list(meshes)      # the table holds raw device pointers: keep their owners alive with it
    return table
