// host.cu -- host classes behind include/beam_host.hpp.  They own device memory and call the extern "C" launchers of
// include/beam_b200.h Part 1 exactly where the reference's host classes call theirs:
//   SceneTree::updateGPUScene -> bmResetScene, bmInsertMeshInTree per mesh (SceneTree.cpp:70-107) [+ bmBuildTree]
//   SceneTree::march          -> bmMarch (SceneTree.cpp:109-127)
//   Camera::clear             -> bmClear (Camera.cpp:74-83)
// plus a flat C view of the classes (beam*) so tests and bench.py can drive them through ctypes with HOST buffers.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include <algorithm>
#include "bm_device.cuh"
#include "../../include/beam_host.hpp"

namespace Beam
{
// All work of the host classes -- uploads, kernels, clears -- is issued on ONE non-blocking stream per device (not the legacy default
// stream, which would serialise against every other stream of the process); read-backs run on a second stream behind an event of
// the frame they read, so the download of frame i overlaps the kernels of frame i+1 (two render targets = double buffering).
struct HostStreams { cudaStream_t compute = nullptr, copy = nullptr; };
static HostStreams& host_streams()
{
    static HostStreams per[bm::MAX_DEVICES];
    static std::mutex m;
    int dev = bm::current_device();
    std::lock_guard<std::mutex> lk(m);
    HostStreams& h = per[dev];
    if (!h.compute)
    {
        if (cudaStreamCreateWithFlags(&h.compute, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); h.compute = nullptr; }
        if (cudaStreamCreateWithFlags(&h.copy, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); h.copy = nullptr; }
        bmSetPart1Stream(h.compute);         // bmMarch / bmClear of this device follow
    }
    return h;
}
static cudaStream_t host_stream() { return host_streams().compute; }

// RAII device allocation (DeviceBuffer.h/.cpp with size_t sizes)
class DeviceBuffer
{
public:
    explicit DeviceBuffer(size_t size) : m_size(size) { if (cudaMalloc(&m_dev, size ? size : 1) != cudaSuccess) { cudaGetLastError(); m_dev = nullptr; } }
    ~DeviceBuffer() { if (m_dev) cudaFree(m_dev); }
    bool ok() const { return m_dev != nullptr; }
    void copyFrom(const void* host, bool wait)
    {
        cudaMemcpyAsync(m_dev, host, m_size, cudaMemcpyHostToDevice, host_stream());
        if (wait) cudaStreamSynchronize(host_stream());
    }
    void copyTo(void* host, bool wait) const
    {
        cudaMemcpyAsync(host, m_dev, m_size, cudaMemcpyDeviceToHost, host_stream());
        if (wait) cudaStreamSynchronize(host_stream());
    }
    template <typename T> T* ptr() const { return reinterpret_cast<T*>(m_dev); }
    size_t size() const { return m_size; }
private:
    void* m_dev = nullptr; size_t m_size;
};

class RenderTarget : public IRenderTarget
{
public:
    RenderTarget(u32 w, u32 h, u32 pitch) : m_pitch(pitch), m_width(w), m_height(h), m_store((size_t)pitch * h)
    {
        if (cudaEventCreateWithFlags(&m_written, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); m_written = nullptr; }
    }
    ~RenderTarget() { if (m_locked) unlock(); if (m_written) cudaEventDestroy(m_written); }
    void* buffer() const override { return m_locked ? m_store.ptr<void>() : nullptr; }
    u32 pitch() const override { return m_pitch; }
    u32 width() const override { return m_width; }
    u32 height() const override { return m_height; }
    u32 lock() override { if (m_locked) return ERROR_UNLOCK_FIRST; m_locked = true; s_current = this; return ERROR_ALL_FINE; }
    u32 unlock() override { if (!m_locked) return ERROR_LOCK_FIRST; m_locked = false; if (s_current == this) s_current = nullptr; return ERROR_ALL_FINE; }
    // every launch that writes this target records the event: a read-back waits for exactly the work that produced its pixels
    void markWritten() { if (m_written) cudaEventRecord(m_written, host_stream()); }
    u32 readPixels(u32* host) const override
    {
        if (!host) return ERROR_INVALID_PARAMETER;
        // the trace kernels write row-major without pitch, like bmMarchKernel (BuildTree.cu:489-495 indexes buffer[i]).
        // Copy stream + event: with a second render target in flight the copy overlaps the next frame's kernels; `host` should be
        // pinned (beamHostAlloc) for the copy to be a single DMA.
        cudaStream_t cs = host_streams().copy;
        if (m_written && cudaStreamWaitEvent(cs, m_written, 0) != cudaSuccess) return BM_ERROR_CUDA;
        if (cudaMemcpyAsync(host, m_store.ptr<void>(), (size_t)m_width * m_height * 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return BM_ERROR_CUDA;
        if (cudaStreamSynchronize(cs) != cudaSuccess) return BM_ERROR_CUDA;
        int sticky = bmStickyError();
        return sticky ? (u32)sticky : ERROR_ALL_FINE;
    }
    bool ok() const { return m_store.ok(); }
    static IRenderTarget* get() { return s_current; }
private:
    u32 m_pitch, m_width, m_height; bool m_locked = false; DeviceBuffer m_store;
    cudaEvent_t m_written = nullptr;
    static IRenderTarget* s_current;
};
IRenderTarget* RenderTarget::s_current = nullptr;

sptr<IRenderTarget> IRenderTarget::createHeadless(u32 width, u32 height, u32 pitch)
{
    if (width == 0 || height == 0 || pitch < width * 4) return nullptr;        // argument rule of RenderTarget.cpp:19
    auto rt = std::make_shared<RenderTarget>(width, height, pitch);
    if (!rt->ok()) return nullptr;
    return rt;
}

class Scene;
static void scene_mesh_changed(const wptr<IScene>& scene);      // defined after Scene
class Mesh : public IMesh
{
public:
    Mesh() { for (u32& s : m_sizes) s = 0; }
    // argument rules of Mesh.cpp:30-54
    u32 setVertexData(const float* data, u32 numVertices, u32 numComponents, u32 slotId, bool asyncCopy) override
    {
        if (!data || numVertices == 0 || slotId >= VERTEX_DATA_COUNT || numComponents > 4 || (m_numVertices != 0 && m_numVertices != numVertices) ||
            (slotId == VERTEX_DATA_POSITION && numComponents != 3))
            return ERROR_INVALID_PARAMETER;
        auto buf = std::make_shared<DeviceBuffer>(sizeof(float) * (size_t)numComponents * numVertices);
        if (!buf->ok()) return ERROR_GPU_ALLOC_FAIL;
        buf->copyFrom(data, !asyncCopy);
        m_data[slotId] = buf; m_sizes[slotId] = numComponents; m_numVertices = numVertices;
        scene_mesh_changed(m_scene);          // the owning scene's pointer tables name the replaced buffer
        return ERROR_ALL_FINE;
    }
    u32 setIndices(const u32* indices, u32 numIndices, bool asyncCopy) override
    {
        if (!indices || (numIndices % 3) != 0 || numIndices == 0) return ERROR_INVALID_PARAMETER;
        auto buf = std::make_shared<DeviceBuffer>((size_t)numIndices * sizeof(u32));
        if (!buf->ok()) return ERROR_GPU_ALLOC_FAIL;
        buf->copyFrom(indices, !asyncCopy);
        m_indices = buf; m_numIndices = numIndices;
        scene_mesh_changed(m_scene);
        return ERROR_ALL_FINE;
    }
    // the material of the mesh (SharedTypes.h:20-23 bmMaterial{int texture}): width*height host texels 0x00RRGGBB, row 0 = top
    u32 setTexture(const u32* texels, u32 width, u32 height) override
    {
        if (!texels || width == 0 || height == 0) return ERROR_INVALID_PARAMETER;
        auto buf = std::make_shared<DeviceBuffer>((size_t)width * height * sizeof(u32));
        if (!buf->ok()) return ERROR_GPU_ALLOC_FAIL;
        buf->copyFrom(texels, true);
        m_texture = buf; m_texW = width; m_texH = height;
        scene_mesh_changed(m_scene);          // a texture set after addMesh / updateGPUScene counts from the next frame on
        return ERROR_ALL_FINE;
    }
    wptr<IScene> scene() const override { return m_scene; }
    wptr<IScene> m_scene;
    sptr<DeviceBuffer> m_texture;
    u32 m_texW = 0, m_texH = 0;
    u32 m_numVertices = 0, m_numIndices = 0;
    sptr<DeviceBuffer> m_indices, m_data[VERTEX_DATA_COUNT];
    u32 m_sizes[VERTEX_DATA_COUNT];
};
sptr<IMesh> IMesh::create() { return std::make_shared<Mesh>(); }

// Scene + SceneTree (Scene.cpp, SceneTree.cpp).  TREE is the only scene type with behaviour in the reference.
class Scene : public IScene, public std::enable_shared_from_this<Scene>
{
public:
    Scene()
    {
        m_min = { -30.f, -30.f, -30.f }; m_max = { 30.f, 30.f, 30.f };               // SceneTree.cpp:43-45
        m_root = std::make_shared<DeviceBuffer>(bmGetNodeSize());                       // identifies the tree to the bm* launchers
    }
    // the tree behind m_root goes with the scene (the reference frees its pools in ~SceneTree, SceneTree.cpp:55-68): a recycled root
    // address must never find a stale tree
    ~Scene() override { if (m_root && m_root->ok()) bmReleaseScene(m_root->ptr<void>()); }
    void addMesh(const sptr<IMesh>& mesh) override
    {
        auto m = std::static_pointer_cast<Mesh>(mesh);
        m->m_scene = weak_from_this();
        m_meshes.push_back(m); m_dirtyTable = true;
    }
    void removeMesh(const IMesh& mesh) override
    {
        for (auto it = m_meshes.begin(); it != m_meshes.end(); ++it) if (it->get() == &mesh) { (*it)->m_scene.reset(); m_meshes.erase(it); m_dirtyTable = true; break; }
    }
    void meshChanged() { m_dirtyTable = true; }
    // Scene::updateMeshPtrs, Scene.cpp:58-79
    void updateMeshPtrs()
    {
        if (!m_dirtyTable) return;
        std::vector<bmStaticMeshData> tab(m_meshes.size());
        for (size_t i = 0; i < m_meshes.size(); i++)
        {
            auto& m = m_meshes[i];
            memset(&tab[i], 0, sizeof(tab[i]));
            tab[i].m_indices = m->m_indices ? m->m_indices->ptr<u32>() : nullptr;
            for (u32 s = 0; s < VERTEX_DATA_COUNT; s++) { tab[i].m_vertexData[s] = m->m_data[s] ? m->m_data[s]->ptr<float>() : nullptr; tab[i].m_vertexDataSizes[s] = m->m_sizes[s]; }
            tab[i].m_numIndices = m->m_numIndices; tab[i].m_numVertices = m->m_numVertices;
        }
        m_table = std::make_shared<DeviceBuffer>(std::max<size_t>(1, tab.size()) * sizeof(bmStaticMeshData));
        if (!tab.empty()) m_table->copyFrom(tab.data(), true);
        // materials: one bmMaterial per mesh, one texture-table entry per textured mesh
        std::vector<bmMaterial> mats(m_meshes.size());
        std::vector<bmTexture> texs;
        for (size_t i = 0; i < m_meshes.size(); i++)
        {
            auto& m = m_meshes[i];
            mats[i].texture = -1;
            if (m->m_texture) { mats[i].texture = (int)texs.size(); texs.push_back({ m->m_texture->ptr<const u32>(), m->m_texW, m->m_texH }); }
        }
        m_materials = std::make_shared<DeviceBuffer>(std::max<size_t>(1, mats.size()) * sizeof(bmMaterial));
        if (!mats.empty()) m_materials->copyFrom(mats.data(), true);
        m_textures = std::make_shared<DeviceBuffer>(std::max<size_t>(1, texs.size()) * sizeof(bmTexture));
        if (!texs.empty()) m_textures->copyFrom(texs.data(), true);
        m_numTextures = (u32)texs.size();
        m_dirtyTable = false;
    }
    void updateGPUScene() override
    {
        updateMeshPtrs();
        bmResetScene(m_root->ptr<void>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0);
        for (u32 i = 0; i < m_meshes.size(); i++)
        {
            auto& m = m_meshes[i];
            if (!m->m_data[VERTEX_DATA_POSITION] || !m->m_indices) continue;
            bmInsertMeshInTree(m->m_data[VERTEX_DATA_POSITION]->ptr<const bmVec3>(), m->m_indices->ptr<u32>(), m->m_numIndices, i, &m_min, &m_max,
                               m_root->ptr<void>(), nullptr, nullptr, nullptr, nullptr);
        }
        if (bmTree* t = bmTreeFromRoot(m_root->ptr<void>())) bmTreeSetStream(t, host_stream());
        bmBuildTree(m_root->ptr<void>());
    }
    // Scene::march + SceneTree::march, Scene.cpp:81-97, SceneTree.cpp:109-127
    u32 march(const float* eye3, const float* orient9, const sptr<DeviceBuffer>& rays, u32 raysWide, u32 raysHigh)
    {
        IRenderTarget* rt = RenderTarget::get();
        if (!rt) return ERROR_NO_RENDER_TARGET;
        if (rt->width() != raysWide || rt->height() != raysHigh) return ERROR_RT_CAM_MISMATCH;
        updateMeshPtrs();                        // a vertex / index / texture buffer replaced since the last frame: never shade through a freed pointer
        bmVec3 eye = { eye3[0], eye3[1], eye3[2] }; bmMat3 orient; memcpy(orient.m, orient9, 36);
        bmMarch(m_root->ptr<void>(), &m_min, &m_max, rays->ptr<const bmVec3>(), rt->buffer<u32>(), rt->pitch(), rt->width(), rt->height(), &eye, &orient,
                m_table ? m_table->ptr<const bmStaticMeshData>() : nullptr);
        static_cast<RenderTarget*>(rt)->markWritten();
        int sticky = bmStickyError();              // the reference's launchers return void: a failure inside is kept, not fatal
        return sticky ? (u32)sticky : ERROR_ALL_FINE;
    }
    u32 marchHits(const float* eye3, const float* orient9, const sptr<DeviceBuffer>& rays, u32 raysWide, u32 raysHigh, bmHit* hostHits)
    {
        bmTree* t = bmTreeFromRoot(m_root->ptr<void>());
        size_t n = (size_t)raysWide * raysHigh;
        IRenderTarget* rt = RenderTarget::get();
        if (rt && (rt->width() != raysWide || rt->height() != raysHigh)) return ERROR_RT_CAM_MISMATCH;
        if (!t) { for (size_t i = 0; i < n; i++) hostHits[i] = { FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu }; return ERROR_ALL_FINE; }
        updateMeshPtrs();
        DeviceBuffer* hits = scratch(0, n * sizeof(bmHit));
        if (!hits) return ERROR_GPU_ALLOC_FAIL;
        int r = bmTraceCamera(t, rays->ptr<const float>(), n, raysWide, eye3, orient9, BM_TRACE_REFERENCE, hits->ptr<bmHit>(), rt ? rt->buffer<u32>() : nullptr,
                              m_table ? m_table->ptr<const bmStaticMeshData>() : nullptr, nullptr);
        if (r != BM_OK) return (u32)r;
        if (rt) static_cast<RenderTarget*>(rt)->markWritten();
        if (cudaMemcpyAsync(hostHits, hits->ptr<void>(), n * sizeof(bmHit), cudaMemcpyDeviceToHost, host_stream()) != cudaSuccess) return BM_ERROR_CUDA;
        if (cudaStreamSynchronize(host_stream()) != cudaSuccess) return BM_ERROR_CUDA;
        return ERROR_ALL_FINE;
    }
    // hits -> deferred shading with the meshes' materials (ICamera::traceSceneTextured)
    u32 marchTextured(const float* eye3, const float* orient9, const sptr<DeviceBuffer>& rays, u32 raysWide, u32 raysHigh)
    {
        IRenderTarget* rt = RenderTarget::get();
        if (!rt) return ERROR_NO_RENDER_TARGET;
        if (rt->width() != raysWide || rt->height() != raysHigh) return ERROR_RT_CAM_MISMATCH;
        updateMeshPtrs();                        // a texture set after updateGPUScene still counts (Mesh::setTexture marks the scene dirty)
        size_t n = (size_t)raysWide * raysHigh;
        bmTree* t = bmTreeFromRoot(m_root->ptr<void>());
        if (!t) { bmClear(rt->buffer<u32>(), rt->pitch(), raysWide, raysHigh, 255u << 8); return ERROR_ALL_FINE; }
        DeviceBuffer* hits = scratch(0, n * sizeof(bmHit));
        if (!hits) return ERROR_GPU_ALLOC_FAIL;
        int r = bmTraceCamera(t, rays->ptr<const float>(), n, raysWide, eye3, orient9, BM_TRACE_REFERENCE, hits->ptr<bmHit>(), nullptr, nullptr, nullptr);
        if (!r) r = bmShadeHits(t, hits->ptr<bmHit>(), n, m_table->ptr<const bmStaticMeshData>(), m_materials->ptr<const bmMaterial>(),
                                m_textures->ptr<const bmTexture>(), m_numTextures, rt->buffer<u32>());
        if (r != BM_OK) return (u32)r;
        static_cast<RenderTarget*>(rt)->markWritten();
        return ERROR_ALL_FINE;                    // asynchronous like traceScene: the caller synchronises (readPixels does)
    }
    // progressive frame (see ICamera::traceSceneProgressive): every sample = camera rays (jittered inside the pixel when samples > 1; formed
    // in the trace kernel's registers, never materialised) -> trace in the reference's semantics -> one diffuse bounce ray per hit ->
    // CLOSEST-hit trace of the bounce rays (t >= 0: a bounce ray starts 1e-4 above its source triangle, whose plane the reference's
    // any-sign acceptance would "hit" at t < 0) -> accumulate.  Nothing is read back inside the frame: ray accounting stays on the device.
    u32 marchProgressive(const float* eye3, const float* orient9, const sptr<DeviceBuffer>& rays, u32 raysWide, u32 raysHigh, float jx, float jy, u32 samples,
                         u32 seed, u64* raysTraced)
    {
        IRenderTarget* rt = RenderTarget::get();
        if (!rt) return ERROR_NO_RENDER_TARGET;
        if (rt->width() != raysWide || rt->height() != raysHigh) return ERROR_RT_CAM_MISMATCH;
        if (samples == 0) return ERROR_INVALID_PARAMETER;
        size_t n = (size_t)raysWide * raysHigh;
        bmTree* t = bmTreeFromRoot(m_root->ptr<void>());
        if (!t) { bmClear(rt->buffer<u32>(), rt->pitch(), raysWide, raysHigh, 255u << 8); static_cast<RenderTarget*>(rt)->markWritten(); if (raysTraced) *raysTraced = 0; return ERROR_ALL_FINE; }
        cudaStream_t st = host_stream();
        // up to `group` samples of every pixel per launch (bmTraceCameraSamples), then ONE bounce launch for the hits of those samples
        // (bmTraceBounceCamera: the bounce rays are formed in the kernel and accumulated where they end; no ray list, no compaction).  A
        // single sample of a small scene yields too few bounce rays to fill the GPU: f16 at 1080p, 71 K per sample.
        const u32 group = (u32)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(samples, 8), (size_t)(32u << 20) / std::max<size_t>(n, 1)));
        DeviceBuffer *ph1 = scratch(0, n * group * sizeof(bmHit)), *pcnt = scratch(5, 16), *pacc = scratch(6, n * 4), *pcov = scratch(7, n * 4);
        if (!ph1 || !pcnt || !pacc || !pcov) return ERROR_GPU_ALLOC_FAIL;
        DeviceBuffer &h1 = *ph1, &cnt = *pcnt, &acc = *pacc, &cov = *pcov;
        u64* total = reinterpret_cast<u64*>(cnt.ptr<char>() + 8);
        cudaMemsetAsync(acc.ptr<void>(), 0, n * 4, st); cudaMemsetAsync(cov.ptr<void>(), 0, n * 4, st); cudaMemsetAsync(cnt.ptr<void>(), 0, 16, st);
        for (u32 s0 = 0; s0 < samples; s0 += group)
        {
            const u32 g = std::min(group, samples - s0);
            const u32 camSample = samples > 1 ? s0 : 0xFFFFFFFFu;             // one sample per pixel = the exact pixel-centre ray bmMarch traces
            int r = samples > 1 ? bmTraceCameraSamples(t, rays->ptr<const float>(), n, raysWide, eye3, orient9, jx, jy, s0, g, seed, BM_TRACE_REFERENCE, h1.ptr<bmHit>(), nullptr)
                                : bmTraceCameraJittered(t, rays->ptr<const float>(), n, raysWide, eye3, orient9, jx, jy, camSample, seed, BM_TRACE_REFERENCE, h1.ptr<bmHit>(), nullptr, nullptr, nullptr);
            if (!r) r = bmTraceBounceCamera(t, rays->ptr<const float>(), n, raysWide, eye3, orient9, jx, jy, camSample, g, seed, h1.ptr<bmHit>(), s0, 1, seed, nullptr, 0.25f, 0.75f,
                                            acc.ptr<float>(), cov.ptr<u32>(), raysTraced ? total : nullptr, nullptr);
            if (!r && raysTraced) r = bm::add_count(nullptr, n * g, total, st);
            if (r) return (u32)r;
        }
        int r = bmProgressiveResolve(acc.ptr<float>(), cov.ptr<u32>(), n, samples, rt->buffer<u32>(), st);
        if (r) return (u32)r;
        static_cast<RenderTarget*>(rt)->markWritten();
        if (raysTraced)
        {
            // ONE 8-byte read-back per frame, queued BEFORE the render target's "written" event and not waited for: like the pixels, the
            // count is valid once readPixels of this frame's target (or a device synchronise) has returned.  With pinned memory
            // (beamHostAlloc) the call stays asynchronous, so frames can be pipelined; a pageable destination makes the copy block.
            if (cudaMemcpyAsync(raysTraced, total, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return BM_ERROR_CUDA;
            static_cast<RenderTarget*>(rt)->markWritten();
        }
        return ERROR_ALL_FINE;
    }
    void* rootKey() const { return m_root->ptr<void>(); }
private:
    // per-frame scratch (hit records, ray batches, accumulators) is kept between frames and only grows: a frame performs no cudaMalloc
    DeviceBuffer* scratch(int slot, size_t bytes)
    {
        sptr<DeviceBuffer>& b = m_scratch[slot];
        if (!b || b->size() < bytes) { b.reset(); b = std::make_shared<DeviceBuffer>(bytes); if (!b->ok()) { b.reset(); return nullptr; } }
        return b.get();
    }
    sptr<DeviceBuffer> m_scratch[8];
    bmVec3 m_min, m_max;
    std::vector<sptr<Mesh>> m_meshes;
    sptr<DeviceBuffer> m_root, m_table, m_materials, m_textures;
    u32 m_numTextures = 0;
    bool m_dirtyTable = true;
};
sptr<IScene> IScene::create() { return std::make_shared<Scene>(); }
static void scene_mesh_changed(const wptr<IScene>& scene) { if (auto s = scene.lock()) static_cast<Scene*>(s.get())->meshChanged(); }

class Camera : public ICamera
{
public:
    // Camera::setInitialRays, Camera.cpp:43-72: per-pixel unit directions built on the host with accumulating rx / ry
    u32 setInitialRays(u32 width, u32 height, float left, float right, float top, float bottom, float zoom) override
    {
        if (width == 0 || height == 0) return ERROR_INVALID_PARAMETER;
        m_width = width; m_height = height;
        m_pixelW = fabsf((right - left) / width) / zoom; m_pixelH = fabsf((bottom - top) / height) / zoom;     // pixel size in units of zoom (jitter range)
        std::vector<float> tab((size_t)width * height * 3);
        float dx = (right - left) / width;
        float dy = (bottom - top) / height;
        float ry = top + dy * .5f;
        float z2 = zoom * zoom;
        for (u32 y = 0; y < height; y++, ry += dy)
        {
            float rx = left + dx * .5f;
            for (u32 x = 0; x < width; x++, rx += dx)
            {
                float d = 1.f / sqrtf(z2 + rx * rx + ry * ry);
                if (std::isnan(d) || d <= 0.f) return ERROR_INVALID_PARAMETER;
                size_t a = ((size_t)y * width + x) * 3;
                tab[a] = rx * d; tab[a + 1] = ry * d; tab[a + 2] = zoom * d;
            }
        }
        m_rays = std::make_shared<DeviceBuffer>(tab.size() * sizeof(float));
        if (!m_rays->ok()) return ERROR_GPU_ALLOC_FAIL;
        m_rays->copyFrom(tab.data(), true);
        return ERROR_ALL_FINE;
    }
    u32 clear(u32 value) override
    {
        IRenderTarget* rt = RenderTarget::get();
        if (!rt) return ERROR_NO_RENDER_TARGET;
        host_streams();                           // makes the Part-1 launchers of this device use the host classes' stream
        bmClear(rt->buffer<u32>(), rt->pitch(), rt->width(), rt->height(), value);
        static_cast<RenderTarget*>(rt)->markWritten();
        return ERROR_ALL_FINE;
    }
    u32 traceScene(const float* eye3, const float* orient3x3, sptr<IScene>& scene) override
    {
        if (!eye3 || !orient3x3 || !scene || m_width == 0 || m_height == 0 || !m_rays) return ERROR_INVALID_PARAMETER;   // Camera.cpp:88-92
        return static_cast<Scene*>(scene.get())->march(eye3, orient3x3, m_rays, m_width, m_height);
    }
    u32 traceSceneHits(const float* eye3, const float* orient3x3, sptr<IScene>& scene, bmHit* hostHits) override
    {
        if (!eye3 || !orient3x3 || !scene || m_width == 0 || m_height == 0 || !m_rays || !hostHits) return ERROR_INVALID_PARAMETER;
        return static_cast<Scene*>(scene.get())->marchHits(eye3, orient3x3, m_rays, m_width, m_height, hostHits);
    }
    u32 traceSceneProgressive(const float* eye3, const float* orient3x3, sptr<IScene>& scene, u32 samples, u32 seed, u64* raysTraced) override
    {
        if (!eye3 || !orient3x3 || !scene || m_width == 0 || m_height == 0 || !m_rays) return ERROR_INVALID_PARAMETER;
        return static_cast<Scene*>(scene.get())->marchProgressive(eye3, orient3x3, m_rays, m_width, m_height, m_pixelW, m_pixelH, samples, seed, raysTraced);
    }
    u32 traceSceneTextured(const float* eye3, const float* orient3x3, sptr<IScene>& scene) override
    {
        if (!eye3 || !orient3x3 || !scene || m_width == 0 || m_height == 0 || !m_rays) return ERROR_INVALID_PARAMETER;
        return static_cast<Scene*>(scene.get())->marchTextured(eye3, orient3x3, m_rays, m_width, m_height);
    }
    u32 copyInitialRays(float* host) const { if (!m_rays) return ERROR_INVALID_PARAMETER; m_rays->copyTo(host, true); return ERROR_ALL_FINE; }
private:
    u32 m_width = 0, m_height = 0;
    float m_pixelW = 0.f, m_pixelH = 0.f;
    sptr<DeviceBuffer> m_rays;
};
sptr<ICamera> ICamera::create() { return std::make_shared<Camera>(); }
} // namespace Beam

// ---- flat C view for ctypes (handles are heap-allocated shared_ptr copies) -------------------------------------------------
using namespace Beam;
extern "C"
{
void* beamSceneCreate() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); bm::set_error("no usable CUDA device: this library has no CPU path"); return nullptr; } return new sptr<IScene>(IScene::create()); }
void  beamSceneDestroy(void* s) { delete (sptr<IScene>*)s; }
void* beamMeshCreate() { return new sptr<IMesh>(IMesh::create()); }
void  beamMeshDestroy(void* m) { delete (sptr<IMesh>*)m; }
uint32_t beamMeshSetVertexData(void* m, const float* hostData, uint32_t numVertices, uint32_t numComponents, uint32_t slotId, int asyncCopy)
{ return (*(sptr<IMesh>*)m)->setVertexData(hostData, numVertices, numComponents, slotId, asyncCopy != 0); }
uint32_t beamMeshSetIndices(void* m, const uint32_t* hostIndices, uint32_t numIndices, int asyncCopy)
{ return (*(sptr<IMesh>*)m)->setIndices(hostIndices, numIndices, asyncCopy != 0); }
uint32_t beamMeshSetTexture(void* m, const uint32_t* hostTexels, uint32_t width, uint32_t height) { return (*(sptr<IMesh>*)m)->setTexture(hostTexels, width, height); }
uint32_t beamCameraTraceSceneTextured(void* c, const float* eye3, const float* orient9, void* s) { return (*(sptr<ICamera>*)c)->traceSceneTextured(eye3, orient9, *(sptr<IScene>*)s); }
void  beamSceneAddMesh(void* s, void* m) { (*(sptr<IScene>*)s)->addMesh(*(sptr<IMesh>*)m); }
void  beamSceneRemoveMesh(void* s, void* m) { (*(sptr<IScene>*)s)->removeMesh(**(sptr<IMesh>*)m); }
void  beamSceneUpdateGPUScene(void* s) { (*(sptr<IScene>*)s)->updateGPUScene(); }
void* beamSceneTree(void* s) { return bmTreeFromRoot(static_cast<Scene*>(((sptr<IScene>*)s)->get())->rootKey()); }
void* beamCameraCreate() { return new sptr<ICamera>(ICamera::create()); }
void  beamCameraDestroy(void* c) { delete (sptr<ICamera>*)c; }
uint32_t beamCameraSetInitialRays(void* c, uint32_t w, uint32_t h, float l, float r, float t, float b, float zoom) { return (*(sptr<ICamera>*)c)->setInitialRays(w, h, l, r, t, b, zoom); }
uint32_t beamCameraCopyInitialRays(void* c, float* host) { return static_cast<Camera*>(((sptr<ICamera>*)c)->get())->copyInitialRays(host); }
uint32_t beamCameraClear(void* c, uint32_t value) { return (*(sptr<ICamera>*)c)->clear(value); }
uint32_t beamCameraTraceScene(void* c, const float* eye3, const float* orient9, void* s) { return (*(sptr<ICamera>*)c)->traceScene(eye3, orient9, *(sptr<IScene>*)s); }
uint32_t beamCameraTraceSceneHits(void* c, const float* eye3, const float* orient9, void* s, bmHit* hostHits) { return (*(sptr<ICamera>*)c)->traceSceneHits(eye3, orient9, *(sptr<IScene>*)s, hostHits); }
uint32_t beamCameraTraceSceneProgressive(void* c, const float* eye3, const float* orient9, void* s, uint32_t samples, uint32_t seed, uint64_t* raysTraced)
{ return (*(sptr<ICamera>*)c)->traceSceneProgressive(eye3, orient9, *(sptr<IScene>*)s, samples, seed, raysTraced); }
void* beamRenderTargetCreateHeadless(uint32_t w, uint32_t h, uint32_t pitch) { auto rt = IRenderTarget::createHeadless(w, h, pitch); return rt ? new sptr<IRenderTarget>(rt) : nullptr; }
void  beamRenderTargetDestroy(void* r) { delete (sptr<IRenderTarget>*)r; }
uint32_t beamRenderTargetLock(void* r) { return (*(sptr<IRenderTarget>*)r)->lock(); }
uint32_t beamRenderTargetUnlock(void* r) { return (*(sptr<IRenderTarget>*)r)->unlock(); }
void* beamRenderTargetBuffer(void* r) { return (*(sptr<IRenderTarget>*)r)->buffer(); }
uint32_t beamRenderTargetReadPixels(void* r, uint32_t* hostPixels) { return (*(sptr<IRenderTarget>*)r)->readPixels(hostPixels); }
int   beamDeviceSynchronize() { return cudaDeviceSynchronize() == cudaSuccess ? 0 : BM_ERROR_CUDA; }
// pinned host staging for readPixels / traceSceneHits callers (a pageable destination makes the driver stage the copy in chunks)
void* beamHostAlloc(size_t bytes) { void* p = nullptr; if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void  beamHostFree(void* p) { if (p) cudaFreeHost(p); }
void* beamHostStream() { return host_stream(); }
}
