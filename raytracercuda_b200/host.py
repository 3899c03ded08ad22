"""Python mirror of the reference's public interface (Raytracer/Beam.h:32-72: IRenderTarget / IMesh / IScene / ICamera),
bound to the C++ host classes inside libbeam_b200.so (include/beam_host.hpp, beam_host_c.h).  Method names, argument
meaning and u32 error codes are the reference's; all buffers passed in here are HOST (numpy) arrays."""
from __future__ import annotations

import numpy as np

from . import _lib

ERROR_ALL_FINE, ERROR_NO_VERTICES, ERROR_INVALID_PARAMETER, ERROR_GPU_ALLOC_FAIL, ERROR_INVALID_FORMAT = 0, 1, 2, 3, 4
ERROR_RT_CAM_MISMATCH, ERROR_UNLOCK_FIRST, ERROR_LOCK_FIRST, ERROR_NO_RENDER_TARGET = 5, 6, 7, 8
VERTEX_DATA_POSITION, VERTEX_DATA_NORMAL, VERTEX_DATA_UV1, VERTEX_DATA_UV2, VERTEX_DATA_TANGENT = 0, 1, 2, 3, 4
VERTEX_DATA_BITANGENT, VERTEX_DATA_EXTRA1, VERTEX_DATA_EXTRA2, VERTEX_DATA_EXTRA3, VERTEX_DATA_EXTRA4, VERTEX_DATA_COUNT = 5, 6, 7, 8, 9, 10

HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("tri", np.uint32)])


class IMesh:
    def __init__(self):
        self.L = _lib.load()
        self.h = self.L.beamMeshCreate()

    @staticmethod
    def create() -> "IMesh":
        return IMesh()

    def setVertexData(self, vertices, numVertices, numComponents, slotId, asyncCopy=False) -> int:
        a = None if vertices is None else np.ascontiguousarray(vertices, np.float32)
        return self.L.beamMeshSetVertexData(self.h, None if a is None else a.ctypes.data, int(numVertices), int(numComponents), int(slotId), int(asyncCopy))

    def setIndices(self, indices, numIndices, asyncCopy=False) -> int:
        a = None if indices is None else np.ascontiguousarray(indices, np.uint32)
        return self.L.beamMeshSetIndices(self.h, None if a is None else a.ctypes.data, int(numIndices), int(asyncCopy))

    def setTexture(self, texels, width, height) -> int:
        """IMesh::setTexture: (height, width) uint32 texels 0x00RRGGBB, row 0 = top"""
        a = None if texels is None else np.ascontiguousarray(texels, np.uint32)
        return self.L.beamMeshSetTexture(self.h, None if a is None else a.ctypes.data, int(width), int(height))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.beamMeshDestroy(self.h); self.h = None


class IScene:
    def __init__(self):
        self.L = _lib.load()
        self.h = self.L.beamSceneCreate()
        if not self.h:
            raise _lib.BeamError("IScene::create failed: " + self.L.bmLastErrorString().decode())
        self._meshes = []

    @staticmethod
    def create() -> "IScene":
        return IScene()

    def addMesh(self, mesh: IMesh):
        self._meshes.append(mesh); self.L.beamSceneAddMesh(self.h, mesh.h)

    def removeMesh(self, mesh: IMesh):
        self.L.beamSceneRemoveMesh(self.h, mesh.h)
        self._meshes = [m for m in self._meshes if m is not mesh]

    def updateGPUScene(self):
        self.L.beamSceneUpdateGPUScene(self.h)

    def tree_handle(self):
        return self.L.beamSceneTree(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.beamSceneDestroy(self.h); self.h = None


class IRenderTarget:
    def __init__(self, h, w, hgt, pitch):
        self.L = _lib.load(); self.h = h; self._w = w; self._h = hgt; self._pitch = pitch

    @staticmethod
    def createHeadless(width, height, pitch=None):
        L = _lib.load()
        pitch = width * 4 if pitch is None else pitch
        h = L.beamRenderTargetCreateHeadless(int(width), int(height), int(pitch))
        return IRenderTarget(h, width, height, pitch) if h else None

    def width(self): return self._w
    def height(self): return self._h
    def pitch(self): return self._pitch
    def buffer(self): return self.L.beamRenderTargetBuffer(self.h)
    def lock(self) -> int: return self.L.beamRenderTargetLock(self.h)
    def unlock(self) -> int: return self.L.beamRenderTargetUnlock(self.h)

    def readPixels(self) -> np.ndarray:
        px = np.empty((self._h, self._w), np.uint32)
        err = self.L.beamRenderTargetReadPixels(self.h, px.ctypes.data)
        _lib.check(err, "IRenderTarget::readPixels")
        return px

    def __del__(self):
        if getattr(self, "h", None):
            self.L.beamRenderTargetDestroy(self.h); self.h = None


class ICamera:
    def __init__(self):
        self.L = _lib.load(); self.h = self.L.beamCameraCreate(); self._n = 0

    @staticmethod
    def create() -> "ICamera":
        return ICamera()

    def setInitialRays(self, width, height, left=-1.0, right=1.0, top=1.0, bottom=-1.0, zoom=1.0) -> int:
        err = self.L.beamCameraSetInitialRays(self.h, int(width), int(height), float(left), float(right), float(top), float(bottom), float(zoom))
        if err == 0:
            self._n = int(width) * int(height)
        return err

    def initialRays(self) -> np.ndarray:
        out = np.empty((self._n, 3), np.float32)
        _lib.check(self.L.beamCameraCopyInitialRays(self.h, out.ctypes.data), "copyInitialRays")
        return out

    def clear(self, value) -> int:
        return self.L.beamCameraClear(self.h, int(value))

    @staticmethod
    def _f(a, n):
        if a is None:
            return None
        return np.ascontiguousarray(a, np.float32).reshape(-1)[:n].copy()

    def traceScene(self, eye3, orient3x3, scene: IScene) -> int:
        e = self._f(eye3, 3); o = self._f(orient3x3, 9)
        return self.L.beamCameraTraceScene(self.h, None if e is None else e.ctypes.data, None if o is None else o.ctypes.data, None if scene is None else scene.h)

    def traceSceneTextured(self, eye3, orient3x3, scene: IScene) -> int:
        e = self._f(eye3, 3); o = self._f(orient3x3, 9)
        return self.L.beamCameraTraceSceneTextured(self.h, None if e is None else e.ctypes.data, None if o is None else o.ctypes.data, None if scene is None else scene.h)

    def traceSceneHits(self, eye3, orient3x3, scene: IScene):
        e = self._f(eye3, 3); o = self._f(orient3x3, 9)
        hits = np.empty(self._n, HIT_DTYPE)
        err = self.L.beamCameraTraceSceneHits(self.h, e.ctypes.data, o.ctypes.data, scene.h, hits.ctypes.data)
        return err, hits

    def traceSceneProgressive(self, eye3, orient3x3, scene: IScene, samples: int, seed: int = 0):
        import ctypes as C
        e = self._f(eye3, 3); o = self._f(orient3x3, 9)
        rays = C.c_uint64(0)
        err = self.L.beamCameraTraceSceneProgressive(self.h, e.ctypes.data, o.ctypes.data, scene.h, int(samples), int(seed), C.byref(rays))
        if err == 0:
            device_synchronize()          # the count arrives by an asynchronous copy (valid after readPixels or a synchronise)
        return err, int(rays.value)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.beamCameraDestroy(self.h); self.h = None


def device_synchronize():
    _lib.check(_lib.load().beamDeviceSynchronize(), "cudaDeviceSynchronize")
