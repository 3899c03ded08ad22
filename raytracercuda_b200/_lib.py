"""ctypes binding of include/beam_b200.h.  There is no fallback: a missing or unloadable library raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BEAM_B200_LIB") or os.path.join(HERE, "libbeam_b200.so")      # override: A/B runs of kernel variants

BM_OK = 0
BM_TRACE_REFERENCE = 0
BM_TRACE_CLOSEST = 1
BM_NODES_BINARY, BM_NODES_WIDE, BM_NODES_BOTH = 0, 1, 2
ERROR_NAMES = {0: "ERROR_ALL_FINE", 1: "ERROR_NO_VERTICES", 2: "ERROR_INVALID_PARAMETER", 3: "ERROR_GPU_ALLOC_FAIL", 4: "ERROR_INVALID_FORMAT",
               5: "ERROR_RT_CAM_MISMATCH", 6: "ERROR_UNLOCK_FIRST", 7: "ERROR_LOCK_FIRST", 8: "ERROR_NO_RENDER_TARGET",
               100: "BM_ERROR_CUDA", 101: "BM_ERROR_UNSUPPORTED_WORLD", 102: "BM_ERROR_NOT_BUILT", 103: "BM_ERROR_CAPACITY"}

# every symbol include/beam_b200.h declares (tests/test_abi.py checks the library exports each one)
ABI_SYMBOLS = [
    "bmResetScene", "bmInsertMeshInTree", "bmMarch", "bmGetFaceSize", "bmGetFacePtrSize", "bmGetMaterialSize", "bmGetNodeSize",
    "bmGetFaceStoreSize", "bmGetNodeStoreSize", "bmGetFaceGroupStoreSize", "bmClear",
    "bmResetSpace", "bmInsertMeshInSpace", "bmMarchSpace", "bmGetCellSize", "bmGetMaxHashElements",
    "bmMarchProgressive", "bmGetStoreSize", "bmGetFaceClusterSize", "bmGetBvhNodeSize", "bmGetRaySize", "bmGetRayBoxSize", "bmGetRayFaceSize", "bmRayFaceHitClusterSize",
    "bmLastErrorString", "bmLaunchCount", "bmStickyError", "bmStickyErrorString", "bmClearStickyError", "bmReleaseScene", "bmSetPart1Stream",
    "bmTraceCameraJittered", "bmBounceRaysCameraJittered", "bmTreeCreate", "bmTreeDestroy", "bmTreeSetStream", "bmTreeReset", "bmTreeAddMesh", "bmTreeBuild", "bmBuildTree",
    "bmTreeFromRoot", "bmTreeStats", "bmTreeCopyPairs", "bmTreeCopyTopology", "bmTreeCopyChildBoxes", "bmTreeMeshBases", "bmTreeArraySizes", "bmTreeExport", "bmTreeImport", "bmTreeRoot", "bmTreeView", "bmTreeEnableStageTiming", "bmTreeLastBuildStages", "bmTreeNodeLayout", "bmTreeWorld", "bmTreeCopyWideSlots", "bmTreeSetBricks", "bmBounceRaysCameraAppend", "bmProgressiveAccumulateBounce", "bmTraceCameraSamples", "bmTraceBounceCamera",
    "bmTraceRays", "bmTraceRaysTiled", "bmTraceCamera", "bmHashKeys", "bmCameraRays", "bmBounceRays", "bmBounceRaysCamera", "bmTraceRaysIndirect", "bmProgressiveAccumulate", "bmProgressiveResolve", "bmShadeHits",
    # include/beam_host_c.h
    "beamSceneCreate", "beamSceneDestroy", "beamSceneAddMesh", "beamSceneRemoveMesh", "beamSceneUpdateGPUScene", "beamSceneTree",
    "beamMeshCreate", "beamMeshDestroy", "beamMeshSetVertexData", "beamMeshSetIndices",
    "beamCameraCreate", "beamCameraDestroy", "beamCameraSetInitialRays", "beamCameraCopyInitialRays", "beamCameraClear",
    "beamCameraTraceScene", "beamCameraTraceSceneHits", "beamCameraTraceSceneProgressive",
    "beamRenderTargetCreateHeadless", "beamRenderTargetDestroy", "beamRenderTargetLock", "beamRenderTargetUnlock", "beamRenderTargetBuffer",
    "beamRenderTargetReadPixels", "beamDeviceSynchronize", "beamHostAlloc", "beamHostFree", "beamHostStream",
    "beamObjLoad", "beamObjError", "beamObjFree", "beamObjNumVertices", "beamObjNumMeshes", "beamObjHasNormals", "beamObjPositions", "beamObjNormals",
    "beamObjMeshNumIndices", "beamObjMeshIndices", "beamObjMeshName", "beamSceneLoadObj", "beamWritePPM",
    "beamMeshSetTexture", "beamCameraTraceSceneTextured", "beamObjHasUVs", "beamObjUVs", "beamObjMeshTexture", "beamBmpLoad", "beamImageFree",
]


class BeamError(RuntimeError):
    pass


class bmVec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class bmMat3(C.Structure):
    _fields_ = [("m", C.c_float * 9)]


class bmMaterial(C.Structure):
    _fields_ = [("texture", C.c_int)]


class bmTexture(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class bmStaticMeshData(C.Structure):
    _fields_ = [("m_vertexData", C.c_void_p * 10), ("m_vertexDataSizes", C.c_uint32 * 10), ("m_indices", C.c_void_p),
                ("m_numVertices", C.c_uint32), ("m_numIndices", C.c_uint32)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BeamError(f"{LIB_PATH} is missing: build it with `python -m raytracercuda_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.bmLastErrorString.restype = C.c_char_p
    L.bmLaunchCount.restype = u64
    L.bmTreeCreate.argtypes = [C.POINTER(vp), vp, vp, C.c_float, u32]
    L.bmTreeDestroy.argtypes = [vp]; L.bmTreeDestroy.restype = None
    L.bmTreeSetStream.argtypes = [vp, vp]
    L.bmTreeReset.argtypes = [vp]
    L.bmTreeAddMesh.argtypes = [vp, vp, vp, u32, u32]
    L.bmTreeBuild.argtypes = [vp]
    L.bmBuildTree.argtypes = [vp]
    L.bmTreeFromRoot.argtypes = [vp]; L.bmTreeFromRoot.restype = vp
    L.bmTreeStats.argtypes = [vp, C.POINTER(u64)]
    L.bmTreeCopyPairs.argtypes = [vp, vp, vp]
    L.bmTreeCopyTopology.argtypes = [vp, vp, vp, vp, vp]
    L.bmTreeCopyChildBoxes.argtypes = [vp, vp]
    L.bmTreeMeshBases.argtypes = [vp, C.POINTER(u32), u32]
    L.bmTreeArraySizes.argtypes = [vp, C.POINTER(u64)]
    L.bmTreeExport.argtypes = [vp, vp, vp, vp, vp, vp]
    L.bmTreeImport.argtypes = [C.POINTER(vp), vp, vp, vp, vp, vp, u64, u64, u64, u32, u32, vp, vp, C.c_float, u32]
    L.bmTreeEnableStageTiming.argtypes = [vp, i32]
    L.bmTreeLastBuildStages.argtypes = [vp]; L.bmTreeLastBuildStages.restype = C.c_char_p
    L.bmTreeNodeLayout.argtypes = [vp, C.POINTER(u32)]
    L.bmTreeWorld.argtypes = [vp, vp, vp, C.POINTER(C.c_float), C.POINTER(u32)]
    L.bmTreeCopyWideSlots.argtypes = [vp, vp, vp]
    L.bmTreeSetBricks.argtypes = [vp, i32, i32]
    L.bmTreeRoot.argtypes = [vp, C.POINTER(u32)]
    L.bmTreeView.argtypes = [vp, C.POINTER(vp)]
    L.bmTraceRays.argtypes = [vp, vp, u64, i32, vp, vp]
    L.bmTraceRaysTiled.argtypes = [vp, vp, u64, u32, i32, vp, vp]
    L.bmTraceCamera.argtypes = [vp, vp, u64, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), i32, vp, vp, vp, vp]
    L.bmHashKeys.argtypes = [vp, u64, vp, vp, vp]
    L.bmCameraRays.argtypes = [vp, u64, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, u32, u32, vp, vp]
    L.bmBounceRays.argtypes = [vp, vp, vp, u64, u32, u32, u32, vp, vp, vp]
    L.bmBounceRaysCamera.argtypes = [vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, u64, u32, u32, u32, vp, vp, vp]
    L.bmTraceRaysIndirect.argtypes = [vp, vp, vp, u64, i32, vp, vp]
    L.bmProgressiveAccumulate.argtypes = [vp, u64, vp, vp, vp, C.c_float, vp, vp, vp]
    L.bmProgressiveResolve.argtypes = [vp, vp, u64, u32, vp, vp]
    L.bmResetScene.argtypes = [vp] * 7 + [u32] * 3; L.bmResetScene.restype = None
    L.bmInsertMeshInTree.argtypes = [vp, vp, u32, u32, C.POINTER(bmVec3), C.POINTER(bmVec3), vp, vp, vp, vp, vp]; L.bmInsertMeshInTree.restype = None
    L.bmMarch.argtypes = [vp, C.POINTER(bmVec3), C.POINTER(bmVec3), vp, vp, u32, u32, u32, C.POINTER(bmVec3), C.POINTER(bmMat3), vp]
    L.bmMarch.restype = None
    L.bmClear.argtypes = [vp, u32, u32, u32, u32]; L.bmClear.restype = None
    L.bmResetSpace.argtypes = [vp]; L.bmResetSpace.restype = None
    L.bmInsertMeshInSpace.argtypes = [vp, vp, u32, vp, u32, vp]; L.bmInsertMeshInSpace.restype = None
    L.bmMarchSpace.argtypes = [vp, vp, u32, u32, u32, C.POINTER(bmVec3), C.POINTER(bmMat3), vp, vp]; L.bmMarchSpace.restype = None
    L.bmMarchProgressive.argtypes = [vp, u32] + [vp] * 8; L.bmMarchProgressive.restype = None
    for n in ("bmGetCellSize", "bmGetMaxHashElements", "bmGetStoreSize", "bmGetFaceClusterSize", "bmGetBvhNodeSize", "bmGetRaySize", "bmGetRayBoxSize",
              "bmGetRayFaceSize", "bmRayFaceHitClusterSize"):
        getattr(L, n).restype = u32
    for n in ("bmGetFaceSize", "bmGetFacePtrSize", "bmGetMaterialSize", "bmGetNodeSize", "bmGetFaceStoreSize", "bmGetNodeStoreSize", "bmGetFaceGroupStoreSize"):
        getattr(L, n).restype = u32
    f = C.c_float
    for n in ("beamSceneCreate", "beamMeshCreate", "beamCameraCreate", "beamSceneTree", "beamRenderTargetBuffer"):
        getattr(L, n).restype = vp
    L.beamSceneTree.argtypes = [vp]; L.beamRenderTargetBuffer.argtypes = [vp]
    for n in ("beamSceneDestroy", "beamMeshDestroy", "beamCameraDestroy", "beamRenderTargetDestroy", "beamSceneUpdateGPUScene"):
        getattr(L, n).argtypes = [vp]; getattr(L, n).restype = None
    L.beamSceneAddMesh.argtypes = [vp, vp]; L.beamSceneAddMesh.restype = None
    L.beamSceneRemoveMesh.argtypes = [vp, vp]; L.beamSceneRemoveMesh.restype = None
    L.beamMeshSetVertexData.argtypes = [vp, vp, u32, u32, u32, i32]; L.beamMeshSetVertexData.restype = u32
    L.beamMeshSetIndices.argtypes = [vp, vp, u32, i32]; L.beamMeshSetIndices.restype = u32
    L.beamCameraSetInitialRays.argtypes = [vp, u32, u32, f, f, f, f, f]; L.beamCameraSetInitialRays.restype = u32
    L.beamCameraCopyInitialRays.argtypes = [vp, vp]; L.beamCameraCopyInitialRays.restype = u32
    L.beamCameraClear.argtypes = [vp, u32]; L.beamCameraClear.restype = u32
    L.beamCameraTraceScene.argtypes = [vp, vp, vp, vp]; L.beamCameraTraceScene.restype = u32
    L.beamCameraTraceSceneHits.argtypes = [vp, vp, vp, vp, vp]; L.beamCameraTraceSceneHits.restype = u32
    L.beamCameraTraceSceneProgressive.argtypes = [vp, vp, vp, vp, u32, u32, C.POINTER(u64)]; L.beamCameraTraceSceneProgressive.restype = u32
    L.beamRenderTargetCreateHeadless.argtypes = [u32, u32, u32]; L.beamRenderTargetCreateHeadless.restype = vp
    L.beamRenderTargetLock.argtypes = [vp]; L.beamRenderTargetLock.restype = u32
    L.beamRenderTargetUnlock.argtypes = [vp]; L.beamRenderTargetUnlock.restype = u32
    L.beamRenderTargetReadPixels.argtypes = [vp, vp]; L.beamRenderTargetReadPixels.restype = u32
    L.beamObjLoad.argtypes = [C.c_char_p, i32]; L.beamObjLoad.restype = vp
    L.beamObjError.argtypes = [vp]; L.beamObjError.restype = C.c_char_p
    L.beamObjFree.argtypes = [vp]; L.beamObjFree.restype = None
    for n in ("beamObjNumVertices", "beamObjNumMeshes"):
        getattr(L, n).argtypes = [vp]; getattr(L, n).restype = u32
    L.beamObjHasNormals.argtypes = [vp]
    L.beamObjPositions.argtypes = [vp]; L.beamObjPositions.restype = C.POINTER(C.c_float)
    L.beamObjNormals.argtypes = [vp]; L.beamObjNormals.restype = C.POINTER(C.c_float)
    L.beamObjMeshNumIndices.argtypes = [vp, u32]; L.beamObjMeshNumIndices.restype = u32
    L.beamObjMeshIndices.argtypes = [vp, u32]; L.beamObjMeshIndices.restype = C.POINTER(u32)
    L.beamObjMeshName.argtypes = [vp, u32]; L.beamObjMeshName.restype = C.c_char_p
    L.beamSceneLoadObj.argtypes = [vp, C.c_char_p, i32]
    L.beamWritePPM.argtypes = [C.c_char_p, vp, u32, u32]
    L.beamMeshSetTexture.argtypes = [vp, vp, u32, u32]; L.beamMeshSetTexture.restype = u32
    L.beamCameraTraceSceneTextured.argtypes = [vp, vp, vp, vp]; L.beamCameraTraceSceneTextured.restype = u32
    L.beamObjHasUVs.argtypes = [vp]
    L.beamObjUVs.argtypes = [vp]; L.beamObjUVs.restype = C.POINTER(C.c_float)
    L.beamObjMeshTexture.argtypes = [vp, u32]; L.beamObjMeshTexture.restype = C.c_char_p
    L.beamBmpLoad.argtypes = [C.c_char_p, C.POINTER(u32), C.POINTER(u32)]; L.beamBmpLoad.restype = C.POINTER(u32)
    L.beamImageFree.argtypes = [C.POINTER(u32)]; L.beamImageFree.restype = None
    L.bmShadeHits.argtypes = [vp, vp, u64, vp, vp, vp, u32, vp]
    L.bmStickyError.restype = i32
    L.bmStickyErrorString.restype = C.c_char_p
    L.bmClearStickyError.restype = None
    L.bmReleaseScene.argtypes = [vp]; L.bmReleaseScene.restype = None
    L.bmSetPart1Stream.argtypes = [vp]; L.bmSetPart1Stream.restype = None
    fp = C.POINTER(C.c_float)
    L.bmTraceCameraJittered.argtypes = [vp, vp, u64, u32, fp, fp, f, f, u32, u32, i32, vp, vp, vp, vp]
    L.bmBounceRaysCameraJittered.argtypes = [vp, vp, fp, fp, f, f, u32, u32, vp, u64, u32, u32, u32, vp, vp, vp]
    L.bmBounceRaysCameraAppend.argtypes = [vp, vp, fp, fp, f, f, u32, u32, vp, u64, u32, u32, u32, vp, vp, vp]
    L.bmProgressiveAccumulateBounce.argtypes = [vp, vp, vp, u64, f, vp, vp]
    L.bmTraceCameraSamples.argtypes = [vp, vp, u64, u32, fp, fp, f, f, u32, u32, u32, i32, vp, vp]
    L.bmTraceBounceCamera.argtypes = [vp, vp, u64, u32, fp, fp, f, f, u32, u32, u32, vp, u32, u32, u32, vp, f, f, vp, vp, vp, vp]
    L.beamHostAlloc.argtypes = [C.c_size_t]; L.beamHostAlloc.restype = vp
    L.beamHostFree.argtypes = [vp]; L.beamHostFree.restype = None
    L.beamHostStream.restype = vp
    _lib = L
    return L


def check(code: int, what: str = "") -> None:
    if code != BM_OK:
        msg = load().bmLastErrorString().decode(errors="replace")
        raise BeamError(f"{what or 'beam_b200'} failed: {ERROR_NAMES.get(code, code)} {msg}")
