// TEST INFRASTRUCTURE ONLY (see oracle/README.md). Not part of the product.
//
// Headless stand-in for the one reference host file that cannot build on Linux:
// Raytracer/RenderTarget.cpp (windows.h + glatter + cuda_gl_interop, RenderTarget.cpp:3-6).
// It implements the class declared in the reference's own Raytracer/RenderTarget.h with a
// plain malloc'ed pixel buffer, keeping the lock()/unlock()/get() semantics of
// RenderTarget.cpp:53-91 (lock makes the target the process-global "current" one).
// Compiled ONLY into oracle/_ref/libbeam_ref.so, against the reference headers where they
// lie under /root/reference (never copied into this repository).
#include "RenderTarget.h"
#include <cstdlib>
#include <cstring>
using namespace Beam;

// CUDA 0 (host-emulation oracle): the pixel buffer is host memory.  CUDA 1 (oracle/build_ref_cuda.sh: the reference's host
// classes compiled as they ship, in front of either this repo's kernels or the reference's own): it is device memory.
#if CUDA
static void* rt_alloc(size_t n) { void* p = nullptr; CUDA_CALL(cudaMalloc(&p, n)); CUDA_CALL(cudaMemset(p, 0, n)); return p; }
static void  rt_free(void* p) { cudaFree(p); }
#else
static void* rt_alloc(size_t n) { return calloc(n, 1); }
static void  rt_free(void* p) { free(p); }
#endif

namespace Beam
{
    sptr<IRenderTarget> IRenderTarget::registerGLTBO(u32 tbo, u32 width, u32 height, u32 pitch)
    {
        // same argument rule as RenderTarget.cpp:19
        if ( width==0 || tbo==0 || pitch<width*4 ) return nullptr;
        auto rt = std::make_shared<RenderTarget>(width, height, pitch);
        rt->m_glTbo = tbo;
        return rt;
    }

    RenderTarget::RenderTarget(u32 width, u32 height, u32 pitch):
        m_buffer(nullptr), m_width(width), m_height(height), m_pitch(pitch)
    {
        m_glTbo = 0;
    }

    RenderTarget::~RenderTarget()
    {
        if ( m_buffer ) unlock();
    }

    u32 RenderTarget::lock()
    {
        if ( m_buffer ) return ERROR_UNLOCK_FIRST;
        m_buffer = rt_alloc( (size_t)m_pitch * m_height );
        m_RT = this;
        return ERROR_ALL_FINE;
    }

    u32 RenderTarget::unlock()
    {
        if ( !m_buffer ) return ERROR_LOCK_FIRST;
        rt_free( m_buffer );
        m_buffer = nullptr;
        m_RT = nullptr;
        return ERROR_ALL_FINE;
    }

    IRenderTarget* RenderTarget::get() { return m_RT; }
    IRenderTarget* RenderTarget::m_RT;
}
