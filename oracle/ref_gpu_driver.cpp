// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/README.md). Not part of the product.
//
// C API around the reference's UNMODIFIED host classes compiled in their shipping CUDA-1 mode (Scene.cpp, SceneTree.cpp,
// Mesh.cpp, Camera.cpp, DeviceBuffer.cpp, Util.cpp from /root/reference/Raytracer).  oracle/build_ref_cuda.sh links this
// file twice:
//   oracle/_ref/libbeam_refhost_b200.so : reference host objects + THIS REPO's libbeam_b200.so providing bmResetScene /
//                                         bmInsertMeshInTree / bmMarch / bmGet*Size / bmClear  -> proves the drop-in seam;
//   oracle/_ref/libbeam_refcuda.so      : reference host objects + the reference's own BuildTree.cu / RTClear.cu rebuilt for
//                                         sm_100a -> the "reference CUDA kernels on B200" second baseline.
// Everything goes through the reference's public interface (Beam.h:32-72).
#include <cstdio>
#include <cstring>
#include <chrono>
#include <memory>
#include <vector>
#include <iostream>
#include <unistd.h>
#include <fcntl.h>
#include <cuda_runtime.h>
#include "Beam.h"
#include "RenderTarget.h"

using namespace Beam;

namespace
{
    struct Quiet
    {
        int saved;
        Quiet()  { fflush(stdout); std::cout.flush(); saved = dup(1); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1); close(nul); }
        ~Quiet() { fflush(stdout); std::cout.flush(); dup2(saved, 1); close(saved); }
    };
    struct GScene { sptr<IScene> scene; std::vector<sptr<IMesh>> meshes; sptr<ICamera> cam; sptr<IRenderTarget> rt; u32 w = 0, h = 0; };
    double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C"
{
    void* refgpu_scene_create() { Quiet q; auto* s = new GScene(); s->scene = IScene::create(); return s; }
    void refgpu_scene_destroy(void* h) { auto* s = (GScene*)h; if (s->rt) s->rt->unlock(); delete s; }

    int refgpu_scene_add_mesh(void* h, const float* pos, const float* nrm, u32 numVerts, const u32* idx, u32 numIdx)
    {
        auto* s = (GScene*)h;
        sptr<IMesh> m = IMesh::create();
        s->scene->addMesh(m);
        u32 err = m->setIndices(idx, numIdx); if (err) return (int)err;
        err = m->setVertexData(pos, numVerts, 3, VERTEX_DATA_POSITION); if (err) return (int)err;
        if (nrm) { err = m->setVertexData(nrm, numVerts, 3, VERTEX_DATA_NORMAL); if (err) return (int)err; }
        s->meshes.push_back(m);
        return 0;
    }

    // IScene::updateGPUScene + cudaDeviceSynchronize, like Program::renderFrame's first frame (Program.cpp:291-299). Seconds.
    double refgpu_scene_build(void* h)
    {
        auto* s = (GScene*)h;
        Quiet q;
        cudaDeviceSynchronize();
        double t0 = now();
        s->scene->updateGPUScene();
        cudaDeviceSynchronize();
        return now() - t0;
    }

    int refgpu_camera(void* h, u32 W, u32 H, float left, float right, float top, float bottom, float zoom)
    {
        auto* s = (GScene*)h;
        s->cam = ICamera::create();
        u32 err = s->cam->setInitialRays(W, H, left, right, top, bottom, zoom);
        if (err) return (int)err;
        if (s->rt) s->rt->unlock();
        s->rt = IRenderTarget::registerGLTBO(1, W, H, W * 4);
        if (!s->rt) return (int)ERROR_INVALID_PARAMETER;
        s->w = W; s->h = H;
        return (int)s->rt->lock();
    }

    // ICamera::traceScene + cudaDeviceSynchronize (Program.cpp:324-333), `repeat` times; returns the mean seconds per frame
    double refgpu_trace(void* h, const float* eye3, const float* orient9, int repeat, int* err)
    {
        auto* s = (GScene*)h;
        cudaDeviceSynchronize();
        double t0 = now();
        u32 e = 0;
        for (int i = 0; i < repeat && !e; i++) { e = s->cam->traceScene(eye3, orient9, s->scene); cudaDeviceSynchronize(); }
        double t1 = now();
        if (err) *err = (int)e;
        cudaError_t ce = cudaGetLastError();
        if (ce != cudaSuccess && err && !*err) *err = 1000 + (int)ce;
        return (t1 - t0) / (repeat > 0 ? repeat : 1);
    }

    int refgpu_read_pixels(void* h, u32* pixels)
    {
        auto* s = (GScene*)h;
        if (!s->rt || !s->rt->buffer()) return 1;
        return cudaMemcpy(pixels, s->rt->buffer(), (size_t)s->w * s->h * 4, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
    }
}
