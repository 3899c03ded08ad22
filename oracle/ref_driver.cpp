// TEST INFRASTRUCTURE ONLY (see oracle/README.md). Not part of the product.
//
// Thin C API around the UNMODIFIED reference ("Beam", d1kkop/RaytracerCuda) compiled in its own
// serial host-emulation mode (`#define CUDA 0`, Raytracer/Types.h:8, Raytracer/CudaComon.cuh:19-56).
// build_ref.sh compiles the reference sources where they lie under /root/reference and links them with
// this file and ref_headless_rt.cpp into oracle/_ref/libbeam_ref.so.  Everything here goes through the
// reference's public interface (Beam.h:32-72: IScene / IMesh / ICamera / IRenderTarget); the only thing
// it looks at behind that interface is the finished kd-tree, which it walks read-only to emit the
// canonical (leaf path code, mesh, triangle) list that parity is defined on.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <chrono>
#include <memory>
#include <map>
#include <iostream>
#include <unistd.h>
#include <fcntl.h>

#define private public
#define protected public
#include "SceneTree.h"
#include "Mesh.h"
#include "DeviceBuffer.h"
#include "Camera.h"
#undef private
#undef protected
#include "RenderTarget.h"
#include "BuildTree.cuh"

using namespace Beam;
using namespace glm;

extern "C" int triBoxOverlap(float* boxcenter, float* boxhalfsize, vec3 triverts[3]);

namespace
{
    struct Quiet   // the reference's CPU path prints one line per leaf (BuildTree.cu:307-360)
    {
        int saved;
        Quiet()  { fflush(stdout); std::cout.flush(); saved = dup(1); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1); close(nul); }
        ~Quiet() { fflush(stdout); std::cout.flush(); dup2(saved, 1); close(saved); }
    };

    struct RefScene
    {
        sptr<IScene> scene;
        std::vector<sptr<IMesh>> meshes;
        std::vector<std::vector<u32>> indices;   // host copy, to recover the triangle index of a face record
        // canonical dump
        std::vector<u32> code, mesh, tri;
        u32 numNodes = 0, numLeaves = 0, maxDepth = 0, minLeafDepth = 0, overflowLeaves = 0, maxLeafFaces = 0;
        bool dumped = false;
    };

    double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C"
{
    void* ref_scene_create()
    {
        Quiet q;
        auto* s = new RefScene();
        s->scene = IScene::create();           // Scene.cpp:20-30 -> SceneTree (TREE_TYPE == TREE)
        return s;
    }

    void ref_scene_destroy(void* h) { delete (RefScene*)h; }

    // One IMesh per call, exactly like TestProgram/Model.cpp:52-112 does per aiMesh.
    int ref_scene_add_mesh(void* h, const float* pos, const float* nrm, u32 numVerts, const u32* idx, u32 numIdx)
    {
        auto* s = (RefScene*)h;
        sptr<IMesh> m = IMesh::create();
        s->scene->addMesh(m);
        u32 err = m->setIndices(idx, numIdx);
        if (err) return (int)err;
        err = m->setVertexData(pos, numVerts, 3, VERTEX_DATA_POSITION);
        if (err) return (int)err;
        if (nrm) { err = m->setVertexData(nrm, numVerts, 3, VERTEX_DATA_NORMAL); if (err) return (int)err; }
        s->meshes.push_back(m);
        s->indices.emplace_back(idx, idx + numIdx);
        s->dumped = false;
        return 0;
    }

    // IScene::updateGPUScene (SceneTree.cpp:70-91). Returns seconds spent inside the call.
    double ref_scene_build(void* h)
    {
        auto* s = (RefScene*)h;
        Quiet q;
        double t0 = now();
        s->scene->updateGPUScene();
        double t1 = now();
        s->dumped = false;
        return t1 - t0;
    }

    static void dump(RefScene* s)
    {
        if (s->dumped) return;
        auto* st = static_cast<SceneTree*>(s->scene.get());
        bmTreeNode* root = st->m_rootNode->ptr<bmTreeNode>();
        bmFace* faceBase = st->m_facesBuffer->ptr<bmFace>();
        u32 numRecs = st->m_faceStore->ptr<bmStore<bmFace>>()->m_top;

        // face record k -> triangle index inside its mesh. Records are allocated in (mesh, triangle, leaf) order
        // by the serial path (bmStore::getNew, CudaComon.cuh:268-291), so one forward cursor per mesh suffices.
        std::vector<u32> recTri(numRecs);
        std::vector<u32> cursor(s->indices.size(), 0);
        for (u32 k = 0; k < numRecs; ++k)
        {
            uint4 id = faceBase[k].m_index;
            const auto& ind = s->indices[id.w];
            u32& c = cursor[id.w];
            while (!(ind[3*c] == id.x && ind[3*c+1] == id.y && ind[3*c+2] == id.z)) { ++c; if (3*c >= ind.size()) { fprintf(stderr, "ref_driver: face record without triangle\n"); abort(); } }
            recTri[k] = c;
        }

        s->code.clear(); s->mesh.clear(); s->tri.clear();
        s->numNodes = s->numLeaves = s->maxDepth = s->overflowLeaves = s->maxLeafFaces = 0; s->minLeafDepth = 1000;
        struct E { bmTreeNode* n; u32 depth; u64 path; };
        std::vector<E> stack; stack.push_back({root, 0, 0});
        struct P { u64 path; u32 depth, mesh, tri; };
        std::vector<P> pairs;
        while (!stack.empty())
        {
            E e = stack.back(); stack.pop_back();
            if (e.n != root) s->numNodes++;
            if (!e.n->m_left && !e.n->m_right)
            {
                if (e.n == root && !e.n->m_faces) continue;   // empty scene
                s->numLeaves++;
                s->maxDepth = std::max(s->maxDepth, e.depth);
                s->minLeafDepth = std::min(s->minLeafDepth, e.depth);
                s->maxLeafFaces = std::max(s->maxLeafFaces, e.n->m_faceInsertIdx);
                if (e.n->m_faceInsertIdx > MAX_FACES_PER_BOX) s->overflowLeaves++;
                u32 cnt = std::min<u32>(MAX_FACES_PER_BOX, e.n->m_faceInsertIdx);
                for (u32 i = 0; i < cnt; ++i)
                {
                    bmFace* f = e.n->m_faces[i];
                    pairs.push_back({e.path, e.depth, f->m_index.w, recTri[(u32)(f - faceBase)]});
                }
            }
            else
            {
                if (e.n->m_left)  stack.push_back({e.n->m_left,  e.depth + 1, (e.path << 1)});
                if (e.n->m_right) stack.push_back({e.n->m_right, e.depth + 1, (e.path << 1) | 1});
            }
        }
        std::sort(pairs.begin(), pairs.end(), [](const P& a, const P& b) {
            if (a.path != b.path) return a.path < b.path;
            if (a.mesh != b.mesh) return a.mesh < b.mesh;
            return a.tri < b.tri; });
        for (auto& p : pairs)
        {
            if (p.depth > 32) { fprintf(stderr, "ref_driver: leaf deeper than 32\n"); abort(); }
            s->code.push_back((u32)p.path); s->mesh.push_back(p.mesh); s->tri.push_back(p.tri);
        }
        s->dumped = true;
    }

    // stats[0..6] = pairs, non-root nodes, leaves, max leaf depth, min leaf depth, leaves over 256 faces, max faces in a leaf
    void ref_scene_stats(void* h, uint64_t* stats)
    {
        auto* s = (RefScene*)h; dump(s);
        stats[0] = s->code.size(); stats[1] = s->numNodes; stats[2] = s->numLeaves; stats[3] = s->maxDepth;
        stats[4] = s->minLeafDepth; stats[5] = s->overflowLeaves; stats[6] = s->maxLeafFaces;
    }

    // (leaf path code MSB-first, mesh, triangle), sorted by (code, mesh, triangle). Caller sizes arrays by stats[0].
    void ref_scene_pairs(void* h, u32* code, u32* mesh, u32* tri)
    {
        auto* s = (RefScene*)h; dump(s);
        size_t n = s->code.size();
        memcpy(code, s->code.data(), n * 4); memcpy(mesh, s->mesh.data(), n * 4); memcpy(tri, s->tri.data(), n * 4);
    }

    // ICamera::setInitialRays + ICamera::traceScene (Camera.cpp:43-97) into a headless target. Returns the u32 error
    // code of the reference API; *seconds = time inside traceScene only.
    int ref_trace(void* h, u32 W, u32 H, float left, float right, float top, float bottom, float zoom,
                  const float* eye3, const float* orient9, u32* pixels, double* seconds)
    {
        auto* s = (RefScene*)h;
        sptr<ICamera> cam = ICamera::create();
        u32 err = cam->setInitialRays(W, H, left, right, top, bottom, zoom);
        if (err) return (int)err;
        auto rt = IRenderTarget::registerGLTBO(1, W, H, W * 4);
        if (!rt) return (int)ERROR_INVALID_PARAMETER;
        rt->lock();
        double t0 = now();
        err = cam->traceScene(eye3, orient9, s->scene);
        double t1 = now();
        if (seconds) *seconds = t1 - t0;
        if (!err) memcpy(pixels, rt->buffer(), (size_t)W * H * 4);
        rt->unlock();
        return (int)err;
    }

    // ICamera::setInitialRays only: copies the W*H*3 direction table out (CPU mode DeviceBuffer = malloc).
    int ref_initial_rays(u32 W, u32 H, float left, float right, float top, float bottom, float zoom, float* out)
    {
        sptr<ICamera> cam = ICamera::create();
        u32 err = cam->setInitialRays(W, H, left, right, top, bottom, zoom);
        if (err) return (int)err;
        auto* c = static_cast<Camera*>(cam.get());
        memcpy(out, c->m_initialRays->ptr<float>(), (size_t)W * H * 12);
        return 0;
    }

    // ---- primitive known-answer hooks (CudaComon.cuh:117-187, BoxTriangle.cuh:134-222) ----
    float ref_tri_intersect(const float* o, const float* d, const float* v0, const float* v1, const float* v2, float* u, float* v)
    {
        return bmTriIntersect(vec3(o[0],o[1],o[2]), vec3(d[0],d[1],d[2]), vec3(v0[0],v0[1],v0[2]), vec3(v1[0],v1[1],v1[2]), vec3(v2[0],v2[1],v2[2]), *u, *v);
    }
    float ref_box_ray(const float* bmin, const float* bmax, const float* o, const float* invd)
    {
        return bmBoxRayIntersect(vec3(bmin[0],bmin[1],bmin[2]), vec3(bmax[0],bmax[1],bmax[2]), vec3(o[0],o[1],o[2]), vec3(invd[0],invd[1],invd[2]));
    }
    float ref_box_ray_nozero(const float* bmin, const float* bmax, const float* o, const float* invd)
    {
        return bmBoxRayIntersectNoZero(vec3(bmin[0],bmin[1],bmin[2]), vec3(bmax[0],bmax[1],bmax[2]), vec3(o[0],o[1],o[2]), vec3(invd[0],invd[1],invd[2]));
    }
    int ref_tri_box_overlap(const float* center, const float* half, const float* tri9)
    {
        float c[3] = {center[0],center[1],center[2]}, hs[3] = {half[0],half[1],half[2]};
        vec3 t[3] = { vec3(tri9[0],tri9[1],tri9[2]), vec3(tri9[3],tri9[4],tri9[5]), vec3(tri9[6],tri9[7],tri9[8]) };
        return triBoxOverlap(c, hs, t);
    }
    // struct sizes the reference reports through its own getters (BuildTree.cu:546-555)
    u32 bmGetFaceSize(); u32 bmGetFacePtrSize(); u32 bmGetMaterialSize(); u32 bmGetNodeSize();
    u32 bmGetFaceStoreSize(); u32 bmGetNodeStoreSize(); u32 bmGetFaceGroupStoreSize();
    void ref_struct_sizes(u32* out)
    {
        out[0] = bmGetFaceSize(); out[1] = bmGetFacePtrSize(); out[2] = bmGetMaterialSize(); out[3] = bmGetNodeSize();
        out[4] = bmGetFaceStoreSize(); out[5] = bmGetNodeStoreSize(); out[6] = bmGetFaceGroupStoreSize();
        out[7] = (u32)sizeof(StaticMeshData); out[8] = (u32)sizeof(bmRayFaceHitResult); out[9] = (u32)sizeof(bmStackNode);
    }
}
