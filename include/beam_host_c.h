/* beam_host_c.h -- flat C view of the host classes in beam_host.hpp (Beam::IScene / IMesh / ICamera / IRenderTarget,
 * i.e. the reference's public interface Raytracer/Beam.h:32-72), for callers without a C++ compiler (ctypes, cgo, JNI ...).
 * Handles are opaque; every data pointer here is a HOST pointer -- uploads and read-backs happen inside the calls, exactly
 * where the reference does them (Mesh.cpp:30-54, Camera.cpp:43-72).  Return values are the Beam.h u32 error codes. */
#ifndef BEAM_HOST_C_H
#define BEAM_HOST_C_H
#include "beam_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
void*    beamSceneCreate(void);                                   /* IScene::create(); NULL when no CUDA device is usable */
void     beamSceneDestroy(void* scene);
void     beamSceneAddMesh(void* scene, void* mesh);               /* IScene::addMesh */
void     beamSceneRemoveMesh(void* scene, void* mesh);            /* IScene::removeMesh */
void     beamSceneUpdateGPUScene(void* scene);                    /* IScene::updateGPUScene: builds the tree */
void*    beamSceneTree(void* scene);                              /* the scene's bmTree* (NULL before the first build) */
void*    beamMeshCreate(void);                                    /* IMesh::create() */
void     beamMeshDestroy(void* mesh);
uint32_t beamMeshSetVertexData(void* mesh, const float* hostData, uint32_t numVertices, uint32_t numComponents, uint32_t slotId, int asyncCopy);
uint32_t beamMeshSetIndices(void* mesh, const uint32_t* hostIndices, uint32_t numIndices, int asyncCopy);
uint32_t beamMeshSetTexture(void* mesh, const uint32_t* hostTexels, uint32_t width, uint32_t height);   /* IMesh::setTexture: 0x00RRGGBB, row 0 = top */
void*    beamCameraCreate(void);                                  /* ICamera::create() */
void     beamCameraDestroy(void* camera);
uint32_t beamCameraSetInitialRays(void* camera, uint32_t width, uint32_t height, float left, float right, float top, float bottom, float zoom);
uint32_t beamCameraCopyInitialRays(void* camera, float* host3);   /* the direction table, width*height*3 floats */
uint32_t beamCameraClear(void* camera, uint32_t value);
uint32_t beamCameraTraceScene(void* camera, const float* eye3, const float* orient3x3, void* scene);
uint32_t beamCameraTraceSceneHits(void* camera, const float* eye3, const float* orient3x3, void* scene, bmHit* hostHits);
uint32_t beamCameraTraceSceneTextured(void* camera, const float* eye3, const float* orient3x3, void* scene);   /* ICamera::traceSceneTextured */
uint32_t beamCameraTraceSceneProgressive(void* camera, const float* eye3, const float* orient3x3, void* scene, uint32_t samples, uint32_t seed,
                                         uint64_t* raysTraced /* nullable */);
void*    beamRenderTargetCreateHeadless(uint32_t width, uint32_t height, uint32_t pitch);
void     beamRenderTargetDestroy(void* rt);
uint32_t beamRenderTargetLock(void* rt);
uint32_t beamRenderTargetUnlock(void* rt);
void*    beamRenderTargetBuffer(void* rt);                        /* device pointer while locked */
uint32_t beamRenderTargetReadPixels(void* rt, uint32_t* hostPixels);
int      beamDeviceSynchronize(void);
/* Pinned host staging for read-backs: with a pinned destination IRenderTarget::readPixels is one DMA on the copy stream, and with two
 * render targets the download of frame i overlaps the kernels of frame i+1. */
void*    beamHostAlloc(size_t bytes);
void     beamHostFree(void* pinned);
void*    beamHostStream(void);                                    /* cudaStream_t all host-class work of the current device is issued on */

/* Wavefront OBJ import (replaces TestProgram/Model.cpp + Assimp; rules in raytracercuda_b200/csrc/objload.cpp).  Pure host code. */
void*           beamObjLoad(const char* path, int splitByMaterial);   /* never NULL; check beamObjError */
const char*     beamObjError(void* obj);                              /* NULL when the file parsed */
void            beamObjFree(void* obj);
uint32_t        beamObjNumVertices(void* obj);
uint32_t        beamObjNumMeshes(void* obj);                          /* 1, or one per usemtl group */
int             beamObjHasNormals(void* obj);
const float*    beamObjPositions(void* obj);                          /* numVertices * 3 */
const float*    beamObjNormals(void* obj);                            /* numVertices * 3 or NULL */
uint32_t        beamObjMeshNumIndices(void* obj, uint32_t mesh);
const uint32_t* beamObjMeshIndices(void* obj, uint32_t mesh);
const char*     beamObjMeshName(void* obj, uint32_t mesh);
int             beamObjHasUVs(void* obj);
const float*    beamObjUVs(void* obj);                                /* numVertices * 2 or NULL (`vt`, first use per position wins) */
const char*     beamObjMeshTexture(void* obj, uint32_t mesh);         /* map_Kd path of the mesh's material (relative to the OBJ resolved), or NULL */
/* Uncompressed 24/32-bit BMP -> malloc'ed width*height texels 0x00RRGGBB, row 0 = top (free with beamImageFree); NULL on failure */
uint32_t*       beamBmpLoad(const char* path, uint32_t* width, uint32_t* height);
void            beamImageFree(uint32_t* texels);
/* Model::load: every mesh of the file added to `scene` (positions, normals, UVs when present; with splitByMaterial also the
 * materials' map_Kd BMP textures when they load); returns the number of meshes added, or -1 */
int             beamSceneLoadObj(void* scene, const char* path, int splitByMaterial);
int             beamWritePPM(const char* path, const uint32_t* hostPixels, uint32_t width, uint32_t height);
#ifdef __cplusplus
}
#endif
#endif
