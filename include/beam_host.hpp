// beam_host.hpp -- the reference's public host interface (Raytracer/Beam.h:32-72), kept so that a program written against
// "Beam" switches to libbeam_b200.so without source changes: same namespace, class and method names, argument meaning and
// u32 error codes.  Differences, all forced by the changed subsystems named in BASELINE.json's north_star:
//   * IRenderTarget::createHeadless replaces registerGLTBO (RenderTarget.cpp:17-28): the target is a plain device
//     framebuffer instead of a mapped OpenGL texture-buffer object; lock()/unlock() keep their "make current" semantics
//     (RenderTarget.cpp:53-83);
//   * IRenderTarget::readPixels copies the framebuffer to the host (the viewer's GL present path is out of scope);
//   * buffer sizes are size_t (the reference's u32 DeviceBuffer caps at 4 GiB, DeviceBuffer.h:9);
//   * IMesh::setTexture + ICamera::traceSceneTextured: the UV / material path named by Beam.h:21 and SharedTypes.h:20-23;
//   * ICamera::traceSceneHits also returns hit records (t, u, v, triangle) -- the reference only ever wrote pixels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include "beam_b200.h"

namespace Beam
{
    using u32 = uint32_t;
    using u64 = uint64_t;
    using i32 = int32_t;
    template <typename T> using sptr = std::shared_ptr<T>;
    template <typename T> using wptr = std::weak_ptr<T>;

    constexpr u32 ERROR_ALL_FINE = BM_OK, ERROR_NO_VERTICES = BM_ERROR_NO_VERTICES, ERROR_INVALID_PARAMETER = BM_ERROR_INVALID_PARAMETER,
                  ERROR_GPU_ALLOC_FAIL = BM_ERROR_GPU_ALLOC_FAIL, ERROR_INVALID_FORMAT = BM_ERROR_INVALID_FORMAT,
                  ERROR_RT_CAM_MISMATCH = BM_ERROR_RT_CAM_MISMATCH, ERROR_UNLOCK_FIRST = BM_ERROR_UNLOCK_FIRST,
                  ERROR_LOCK_FIRST = BM_ERROR_LOCK_FIRST, ERROR_NO_RENDER_TARGET = BM_ERROR_NO_RENDER_TARGET;

    constexpr u32 VERTEX_DATA_POSITION = 0, VERTEX_DATA_NORMAL = 1, VERTEX_DATA_UV1 = 2, VERTEX_DATA_UV2 = 3, VERTEX_DATA_TANGENT = 4,
                  VERTEX_DATA_BITANGENT = 5, VERTEX_DATA_EXTRA1 = 6, VERTEX_DATA_EXTRA2 = 7, VERTEX_DATA_EXTRA3 = 8, VERTEX_DATA_EXTRA4 = 9,
                  VERTEX_DATA_COUNT = BM_VERTEX_DATA_COUNT;

    class IScene;

    class IRenderTarget
    {
    public:
        virtual ~IRenderTarget() {}
        static sptr<IRenderTarget> createHeadless(u32 width, u32 height, u32 pitch);   // pitch in bytes, >= width*4
        virtual void* buffer() const = 0;      // device pointer while locked, else nullptr
        virtual u32 pitch() const = 0;
        virtual u32 width() const = 0;
        virtual u32 height() const = 0;
        virtual u32 lock() = 0;
        virtual u32 unlock() = 0;
        virtual u32 readPixels(u32* hostPixels) const = 0;     // width*height words, row-major; synchronises
        template <class T> T* buffer() const { return reinterpret_cast<T*>(buffer()); }
    };

    class IMesh
    {
    public:
        virtual ~IMesh() {}
        static sptr<IMesh> create();
        virtual u32 setVertexData(const float* vertices, u32 numVertices, u32 numComponents, u32 slotId, bool asyncCopy = false) = 0;
        virtual u32 setIndices(const u32* indices, u32 numIndices, bool asyncCopy = false) = 0;
        // the mesh's material texture (bmMaterial{int texture}, SharedTypes.h:20-23, has no setter in the reference): width*height
        // host texels 0x00RRGGBB, row 0 = top of the image; sampled through the UVs in slot VERTEX_DATA_UV1 by traceSceneTextured
        virtual u32 setTexture(const u32* texels, u32 width, u32 height) = 0;
        virtual wptr<IScene> scene() const = 0;
    };

    class IScene
    {
    public:
        virtual ~IScene() {}
        static sptr<IScene> create();
        virtual void addMesh(const sptr<IMesh>& mesh) = 0;
        virtual void removeMesh(const IMesh& mesh) = 0;
        virtual void updateGPUScene() = 0;
    };

    class ICamera
    {
    public:
        virtual ~ICamera() {}
        static sptr<ICamera> create();
        virtual u32 setInitialRays(u32 width, u32 height, float left = -1, float right = 1, float top = 1, float bottom = -1, float zoom = 1) = 0;
        virtual u32 clear(u32 value) = 0;
        virtual u32 traceScene(const float* eye3, const float* orient3x3, sptr<IScene>& scene) = 0;
        // same trace, additionally (or instead: no render target needed when hostHits only) hit records copied to the host
        virtual u32 traceSceneHits(const float* eye3, const float* orient3x3, sptr<IScene>& scene, bmHit* hostHits) = 0;
        // traceScene with the material path the reference declares but never implements: textured meshes show their texture
        // scaled by |n.z|, untextured ones the reference's pixel (bmShadeHits in beam_b200.h states the rule)
        virtual u32 traceSceneTextured(const float* eye3, const float* orient3x3, sptr<IScene>& scene) = 0;
        // what the reference's SceneProgressive (an empty stub, SceneProgressive.cpp:27-77) set out to do: `samples` jittered primary
        // rays per pixel (samples == 1: the exact pixel-centre rays traceScene uses), one diffuse bounce ray per hit traced as a closest hit,
        // accumulated and resolved into the locked render target.  Asynchronous like traceScene; *raysTraced (nullable) is written by an
        // asynchronous copy and is valid once readPixels of that target -- or a device synchronise -- has returned
        virtual u32 traceSceneProgressive(const float* eye3, const float* orient3x3, sptr<IScene>& scene, u32 samples, u32 seed, u64* raysTraced = nullptr) = 0;
    };
}
