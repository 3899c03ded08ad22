// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; the product (raytracercuda_b200/) never links, loads or
// calls this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// Plain C++ restatement (no glm, no CUDA) of the reference's acceleration-structure build and closest-hit traversal,
// each function citing the reference lines it follows (paths relative to /root/reference/Raytracer).
//
// PARITY PINNING: the reference ships no tests or golden vectors.  This restatement is therefore
// pinned against the UNMODIFIED reference compiled in its own CUDA-0 host mode (oracle/build_ref.sh -> oracle/_ref/):
// tests/test_oracle_golden.py asserts identical (leaf code, mesh, triangle) lists, node / leaf counts, pixels and
// primitive results on f16 / suzanne / bunny and random soups, and tests/golden/ holds the fingerprints generated from
// that reference build (tests/golden/make_golden.py), so the pin also holds where /root/reference is absent.
//
// Differences from the reference, all outside what parity is defined on:
//   * dynamic capacity (the reference dies beyond 65,536 leaves / 2^24 face records, SceneTree.cpp:40-41);
//   * returns hit records (t, u, v, mesh, triangle) in addition to the shaded pixel (the reference writes pixels only);
//   * explicit ray batches (origin, direction) besides the camera's primary rays;
//   * meshes without normals shade with the geometric face normal (the reference dereferences a null slot, CudaComon.cuh:259-265);
//   * optional threads over rays (the traversal is read-only; the reference's CPU mode is single-threaded by its globals).
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>

typedef uint32_t u32;
typedef int32_t i32;
typedef uint64_t u64;

namespace
{
struct V3 { float x, y, z; float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
// glm::dot for vec3 = (x*x' + y*y') + z*z'  (glm/detail/func_geometric.inl:54-61)
inline float dot3(V3 a, V3 b) { V3 t = a * b; return t.x + t.y + t.z; }
// glm::cross (func_geometric.inl:74-85)
inline V3 cross3(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
// CudaComon.cuh:71-76 -- NOT fminf/fmaxf: the comparison form decides what a NaN does.
inline float fmin_(float a, float b) { return a < b ? a : b; }
inline float fmax_(float a, float b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------------------------------
// BoxTriangle.cuh:57-79 planeBoxOverlap, :134-222 triBoxOverlap (Akenine-Moller SAT), same operation order.
// ---------------------------------------------------------------------------------------------------------------------
int plane_box_overlap(const float normal[3], const float vert[3], const float maxbox[3])
{
    float vmin[3], vmax[3];
    for (int q = 0; q < 3; q++)
    {
        float v = vert[q];
        if (normal[q] > 0.0f) { vmin[q] = -maxbox[q] - v; vmax[q] = maxbox[q] - v; }
        else                  { vmin[q] = maxbox[q] - v;  vmax[q] = -maxbox[q] - v; }
    }
    if (normal[0] * vmin[0] + normal[1] * vmin[1] + normal[2] * vmin[2] > 0.0f) return 0;
    if (normal[0] * vmax[0] + normal[1] * vmax[1] + normal[2] * vmax[2] >= 0.0f) return 1;
    return 0;
}

// one of the nine edge-cross axes: p = a*v?[i] - b*v?[j] style projections of two vertices, radius from the half size
#define ORC_AXIS(pa, pb, rad)  { float mn, mx; if ((pa) < (pb)) { mn = (pa); mx = (pb); } else { mn = (pb); mx = (pa); } if (mn > (rad) || mx < -(rad)) return 0; }

int tri_box_overlap(const float bc[3], const float hs[3], const V3 tri[3])
{
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3], normal[3];
    for (int k = 0; k < 3; k++) { v0[k] = tri[0][k] - bc[k]; v1[k] = tri[1][k] - bc[k]; v2[k] = tri[2][k] - bc[k]; }
    for (int k = 0; k < 3; k++) { e0[k] = v1[k] - v0[k]; e1[k] = v2[k] - v1[k]; e2[k] = v0[k] - v2[k]; }
    float fex, fey, fez, p0, p1, p2, rad;
    const int X = 0, Y = 1, Z = 2;

    fex = fabsf(e0[X]); fey = fabsf(e0[Y]); fez = fabsf(e0[Z]);
    // AXISTEST_X01(e0[Z], e0[Y], fez, fey)   BoxTriangle.cuh:83-88
    p0 = e0[Z] * v0[Y] - e0[Y] * v0[Z]; p2 = e0[Z] * v2[Y] - e0[Y] * v2[Z]; rad = fez * hs[Y] + fey * hs[Z]; ORC_AXIS(p0, p2, rad)
    // AXISTEST_Y02(e0[Z], e0[X], fez, fex)   :100-105
    p0 = -e0[Z] * v0[X] + e0[X] * v0[Z]; p2 = -e0[Z] * v2[X] + e0[X] * v2[Z]; rad = fez * hs[X] + fex * hs[Z]; ORC_AXIS(p0, p2, rad)
    // AXISTEST_Z12(e0[Y], e0[X], fey, fex)   :119-124
    p1 = e0[Y] * v1[X] - e0[X] * v1[Y]; p2 = e0[Y] * v2[X] - e0[X] * v2[Y]; rad = fey * hs[X] + fex * hs[Y]; ORC_AXIS(p2, p1, rad)

    fex = fabsf(e1[X]); fey = fabsf(e1[Y]); fez = fabsf(e1[Z]);
    p0 = e1[Z] * v0[Y] - e1[Y] * v0[Z]; p2 = e1[Z] * v2[Y] - e1[Y] * v2[Z]; rad = fez * hs[Y] + fey * hs[Z]; ORC_AXIS(p0, p2, rad)
    p0 = -e1[Z] * v0[X] + e1[X] * v0[Z]; p2 = -e1[Z] * v2[X] + e1[X] * v2[Z]; rad = fez * hs[X] + fex * hs[Z]; ORC_AXIS(p0, p2, rad)
    // AXISTEST_Z0(e1[Y], e1[X], fey, fex)    :126-131
    p0 = e1[Y] * v0[X] - e1[X] * v0[Y]; p1 = e1[Y] * v1[X] - e1[X] * v1[Y]; rad = fey * hs[X] + fex * hs[Y]; ORC_AXIS(p0, p1, rad)

    fex = fabsf(e2[X]); fey = fabsf(e2[Y]); fez = fabsf(e2[Z]);
    // AXISTEST_X2(e2[Z], e2[Y], fez, fey)    :90-95
    p0 = e2[Z] * v0[Y] - e2[Y] * v0[Z]; p1 = e2[Z] * v1[Y] - e2[Y] * v1[Z]; rad = fez * hs[Y] + fey * hs[Z]; ORC_AXIS(p0, p1, rad)
    // AXISTEST_Y1(e2[Z], e2[X], fez, fex)    :109-114
    p0 = -e2[Z] * v0[X] + e2[X] * v0[Z]; p1 = -e2[Z] * v1[X] + e2[X] * v1[Z]; rad = fez * hs[X] + fex * hs[Z]; ORC_AXIS(p0, p1, rad)
    p1 = e2[Y] * v1[X] - e2[X] * v1[Y]; p2 = e2[Y] * v2[X] - e2[X] * v2[Y]; rad = fey * hs[X] + fex * hs[Y]; ORC_AXIS(p2, p1, rad)

    // the three box axes (FINDMINMAX, :49-54, :195-208)
    for (int k = 0; k < 3; k++)
    {
        float mn = v0[k], mx = v0[k];
        if (v1[k] < mn) mn = v1[k]; if (v1[k] > mx) mx = v1[k];
        if (v2[k] < mn) mn = v2[k]; if (v2[k] > mx) mx = v2[k];
        if (mn > hs[k] || mx < -hs[k]) return 0;
    }
    // triangle plane (:214-219); CROSS(normal, e0, e1)
    normal[0] = e0[1] * e1[2] - e0[2] * e1[1];
    normal[1] = e0[2] * e1[0] - e0[0] * e1[2];
    normal[2] = e0[0] * e1[1] - e0[1] * e1[0];
    if (!plane_box_overlap(normal, v0, hs)) return 0;
    return 1;
}

// CudaComon.cuh:117-155 bmTriIntersect: Moeller-Trumbore, no epsilon, no culling, no t>0 test, FLT_MAX on miss.
float tri_intersect(V3 orig, V3 dir, V3 v0, V3 v1, V3 v2, float& u, float& v)
{
    V3 v0v1 = v1 - v0;
    V3 v0v2 = v2 - v0;
    V3 pvec = cross3(dir, v0v2);
    float det = dot3(v0v1, pvec);
    float invDet = 1.f / det;
    V3 tvec = orig - v0;
    u = dot3(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return FLT_MAX;
    V3 qvec = cross3(tvec, v0v1);
    v = dot3(dir, qvec) * invDet;
    if (v < 0 || v + u > 1) return FLT_MAX;
    return dot3(v0v2, qvec) * invDet;
}

// CudaComon.cuh:158-172 bmBoxRayIntersect: entry distance clamped to >= 0, FLT_MAX when missed or behind.
float box_ray(V3 bMin, V3 bMax, V3 orig, V3 invDir)
{
    V3 tMin = (bMin - orig) * invDir;
    V3 tMax = (bMax - orig) * invDir;
    float ftmax = fmin_(fmax_(tMin.x, tMax.x), fmin_(fmax_(tMin.y, tMax.y), fmax_(tMin.z, tMax.z)));
    if (ftmax < 0.f) return FLT_MAX;
    float ftmin = fmax_(fmin_(tMin.x, tMax.x), fmax_(fmin_(tMin.y, tMax.y), fmin_(tMin.z, tMax.z)));
    float dist = fmax_(0.f, ftmin);
    return ftmax >= ftmin ? dist : FLT_MAX;
}

// CudaComon.cuh:175-187 bmBoxRayIntersectNoZero
float box_ray_nozero(V3 bMin, V3 bMax, V3 orig, V3 invDir)
{
    V3 tMin = (bMin - orig) * invDir;
    V3 tMax = (bMax - orig) * invDir;
    float ftmin = fmax_(fmin_(tMin.x, tMax.x), fmax_(fmin_(tMin.y, tMax.y), fmin_(tMin.z, tMax.z)));
    float ftmax = fmin_(fmax_(tMin.x, tMax.x), fmin_(fmax_(tMin.y, tMax.y), fmax_(tMin.z, tMax.z)));
    return (std::isinf(ftmin) || ftmin < 0.f) ? ftmax : ftmin;
}

// Hash.cu:17-33 bmHash: Fletcher-16 over the four bytes (little endian) of h
u32 hash32(u32 h)
{
    u32 sum1 = 0, sum2 = 0;
    for (int i = 0; i < 4; i++)
    {
        sum1 = (sum1 + ((h >> (8 * i)) & 0xFF)) % 255;
        sum2 = (sum2 + sum1) % 255;
    }
    return (sum2 << 8) | sum1;
}
// Hash.cu:41-47 bmHash3(uint3)
u32 hash3(u32 x, u32 y, u32 z) { return (hash32(x) + hash32(y) + hash32(z)) % 65536u; }
// Hash.cu:56-59 bmMap; INV_CELL_RES = 1.f/0.03f (Hash.cu:8-9)
i32 map_cell(float f) { return (i32)floorf(f * (1.f / 0.03f)); }

// ---------------------------------------------------------------------------------------------------------------------
// Scene
// ---------------------------------------------------------------------------------------------------------------------
struct Mesh
{
    std::vector<V3> pos; std::vector<V3> nrm; std::vector<u32> idx;
    std::vector<float> uv; u32 uvComps = 0;                       // vertex-data slot 2 (Beam.h:21 VERTEX_DATA_UV1)
    std::vector<u32> texels; u32 texW = 0, texH = 0;              // the mesh's material texture, 0x00RRGGBB, row 0 = top
};
struct FaceRec { u32 i0, i1, i2, mesh, tri; };                    // bmFace (SharedTypes.h:25-29) + the triangle index
struct Node { i32 left, right; i32 leaf; u32 faceInsertIdx; };    // bmTreeNode (BuildTree.cuh:26-36); leaf = index into leafFaces
struct StackNode { V3 mn, mx; i32 node; u32 depth, axis; };       // bmStackNode (BuildTree.cuh:39-48)

struct Scene
{
    V3 wmin, wmax; float minLeaf; u32 maxDepth, maxFacesPerLeaf;
    std::vector<Mesh> meshes;
    std::vector<Node> nodes;                       // nodes[0] = root
    std::vector<FaceRec> faces;                    // one per (triangle, leaf), allocation order (bmStore<bmFace>)
    std::vector<std::vector<u32>> leafFaces;       // per leaf: face record ids in insertion order, capped like the reference
    // the same lists flattened after a build (leaf l = leafFlat[leafFirst[l] .. + leafCount[l])): what the traversal reads
    std::vector<u32> leafFlat, leafCount; std::vector<u64> leafFirst;
    u64 overflowLeaves = 0;
    // canonical dump
    std::vector<u32> pc, pm, pt, bdepth, bprefix;
    u64 nLeaves = 0, maxLeafDepth = 0, minLeafDepth = 0, maxLeafFaces = 0;
};

// BuildTree.cu:154-256 bmInsertTriangleInTree for one triangle (serial form of the kernel body)
void insert_triangle(Scene& S, u32 meshIdx, u32 triIdx)
{
    const Mesh& M = S.meshes[meshIdx];
    u32 i0 = M.idx[3 * triIdx], i1 = M.idx[3 * triIdx + 1], i2 = M.idx[3 * triIdx + 2];
    V3 v[3] = { M.pos[i0], M.pos[i1], M.pos[i2] };
    std::vector<StackNode> stack(S.maxDepth + 2);
    stack[0] = { S.wmin, S.wmax, 0, 0, 0 };
    i32 top = 0;
    do
    {
        StackNode st = stack[top--];
        V3 bSize = st.mx - st.mn;
        float dMin = fmin_(bSize.x, fmin_(bSize.y, bSize.z));
        if ((dMin < S.minLeaf) || (st.depth == S.maxDepth - 1))          // :198-200
        {
            // bmTreeNode::insertFace, BuildTree.cu:36-61
            u32 rec = (u32)S.faces.size();
            S.faces.push_back({ i0, i1, i2, meshIdx, triIdx });
            Node& n = S.nodes[st.node];
            if (n.leaf < 0) { n.leaf = (i32)S.leafFaces.size(); S.leafFaces.emplace_back(); }
            u32 slot = n.faceInsertIdx++;
            if (slot < S.maxFacesPerLeaf) S.leafFaces[n.leaf].push_back(rec);
        }
        else
        {
            u32 ax = st.axis, nAxis = (ax + 1) % 3, ndepth = st.depth + 1;
            float s = .5f * (st.mn[ax] + st.mx[ax]);                      // :216
            V3 lMax = st.mx, rMin = st.mn;
            lMax[ax] = s; rMin[ax] = s;
            V3 bc = (lMax + st.mn) * .5f, hs = (lMax - st.mn) * .5f;      // :221-222
            bool b1 = tri_box_overlap(&bc.x, &hs.x, v) != 0;
            bc = (st.mx + rMin) * .5f; hs = (st.mx - rMin) * .5f;         // :224-225
            bool b2 = tri_box_overlap(&bc.x, &hs.x, v) != 0;
            // bmTreeNode::split2, BuildTree.cu:21-34
            if (b1 && S.nodes[st.node].left < 0)  { S.nodes.push_back({ -1, -1, -1, 0 }); S.nodes[st.node].left = (i32)S.nodes.size() - 1; }
            if (b2 && S.nodes[st.node].right < 0) { S.nodes.push_back({ -1, -1, -1, 0 }); S.nodes[st.node].right = (i32)S.nodes.size() - 1; }
            if (b1) stack[++top] = { st.mn, lMax, S.nodes[st.node].left, ndepth, nAxis };
            if (b2) stack[++top] = { rMin, st.mx, S.nodes[st.node].right, ndepth, nAxis };
        }
    } while (top >= 0);
}

void flatten_leaf_lists(Scene& S)
{
    S.leafFlat.clear(); S.leafFirst.clear(); S.leafCount.clear();
    for (auto& l : S.leafFaces) { S.leafFirst.push_back(S.leafFlat.size()); S.leafCount.push_back((u32)l.size()); S.leafFlat.insert(S.leafFlat.end(), l.begin(), l.end()); }
}

void canonical_dump(Scene& S)
{
    struct P { u64 path; u32 depth, mesh, tri; };
    std::vector<P> pairs;
    S.bdepth.clear(); S.bprefix.clear();
    S.nLeaves = 0; S.maxLeafDepth = 0; S.minLeafDepth = 1000; S.maxLeafFaces = 0; S.overflowLeaves = 0;
    struct E { i32 n; u32 depth; u64 path; };
    std::vector<E> st; st.push_back({ 0, 0, 0 });
    std::vector<std::pair<u32, u32>> branches;
    while (!st.empty())
    {
        E e = st.back(); st.pop_back();
        const Node& n = S.nodes[e.n];
        if (n.left < 0 && n.right < 0)
        {
            if (n.leaf < 0) continue;
            S.nLeaves++;
            S.maxLeafDepth = std::max<u64>(S.maxLeafDepth, e.depth); S.minLeafDepth = std::min<u64>(S.minLeafDepth, e.depth);
            S.maxLeafFaces = std::max<u64>(S.maxLeafFaces, n.faceInsertIdx);
            if (n.faceInsertIdx > S.maxFacesPerLeaf) S.overflowLeaves++;
            for (u64 k = S.leafFirst[n.leaf], ke = k + S.leafCount[n.leaf]; k < ke; k++) { const FaceRec& f = S.faces[S.leafFlat[k]]; pairs.push_back({ e.path, e.depth, f.mesh, f.tri }); }
        }
        else
        {
            if (n.left >= 0 && n.right >= 0) branches.push_back({ e.depth, (u32)e.path });
            // right first: the left subtree is popped first, so leaves come out in ascending path order
            if (n.right >= 0) st.push_back({ n.right, e.depth + 1, (e.path << 1) | 1 });
            if (n.left >= 0)  st.push_back({ n.left, e.depth + 1, e.path << 1 });
        }
    }
    auto less = [](const P& a, const P& b) {
        if (a.path != b.path) return a.path < b.path; if (a.mesh != b.mesh) return a.mesh < b.mesh; return a.tri < b.tri; };
    if (!std::is_sorted(pairs.begin(), pairs.end(), less)) std::sort(pairs.begin(), pairs.end(), less);
    S.pc.clear(); S.pm.clear(); S.pt.clear();
    for (auto& p : pairs) { S.pc.push_back((u32)p.path); S.pm.push_back(p.mesh); S.pt.push_back(p.tri); }
    std::sort(branches.begin(), branches.end());
    for (auto& b : branches) { S.bdepth.push_back(b.first); S.bprefix.push_back(b.second); }
}

struct Hit { float t, u, v; u32 mesh, tri, tie; };

// bmFaceInterpolate<vec3> on slot 1 + normalize + pack, CudaComon.cuh:253-266, BuildTree.cu:486-496
u32 shade(const Scene& S, const Hit& h)
{
    if (h.tri == 0xFFFFFFFFu) return 255u << 8;
    const Mesh& M = S.meshes[h.mesh];
    u32 i0 = M.idx[3 * h.tri], i1 = M.idx[3 * h.tri + 1], i2 = M.idx[3 * h.tri + 2];
    V3 n;
    if (!M.nrm.empty())
    {
        float w = 1 - (h.u + h.v);
        n = M.nrm[i0] * w + M.nrm[i1] * h.u + M.nrm[i2] * h.v;
    }
    else n = cross3(M.pos[i1] - M.pos[i0], M.pos[i2] - M.pos[i0]);      // extension: the reference has no such path
    n = n * (1.f / sqrtf(dot3(n, n)));                                    // glm::normalize = v * inversesqrt(dot(v,v))
    return (u32)fabsf(n.z * 255) << 16;
}

// Material path.  The reference declares it (SharedTypes.h:20-23 bmMaterial{int texture}, Beam.h:21 VERTEX_DATA_UV1, Content/f16.mtl
// map_Kd) but never implements it, so there is nothing of the reference to restate: this is the CPU statement of the rule the product
// defines for bmShadeHits (include/beam_b200.h), kept in the reference's arithmetic style (bmFaceInterpolate weights, no FMA).
u32 shade_textured(const Scene& S, const Hit& h)
{
    if (h.tri == 0xFFFFFFFFu) return 255u << 8;
    const Mesh& M = S.meshes[h.mesh];
    if (M.texels.empty() || M.uv.empty() || M.uvComps < 2) return shade(S, h);
    u32 i0 = M.idx[3 * h.tri], i1 = M.idx[3 * h.tri + 1], i2 = M.idx[3 * h.tri + 2];
    V3 n;
    float w = 1 - (h.u + h.v);
    if (!M.nrm.empty()) n = M.nrm[i0] * w + M.nrm[i1] * h.u + M.nrm[i2] * h.v;
    else n = cross3(M.pos[i1] - M.pos[i0], M.pos[i2] - M.pos[i0]);
    n = n * (1.f / sqrtf(dot3(n, n)));
    u32 c = M.uvComps;
    float tu = M.uv[c * (size_t)i0] * w + M.uv[c * (size_t)i1] * h.u + M.uv[c * (size_t)i2] * h.v;
    float tv = M.uv[c * (size_t)i0 + 1] * w + M.uv[c * (size_t)i1 + 1] * h.u + M.uv[c * (size_t)i2 + 1] * h.v;
    tu = tu - floorf(tu); tv = 1.f - (tv - floorf(tv));
    u32 x = (u32)(tu * (float)M.texW), y = (u32)(tv * (float)M.texH);
    x = x < M.texW ? x : M.texW - 1; y = y < M.texH ? y : M.texH - 1;
    u32 texel = M.texels[(size_t)y * M.texW + x];
    float s = fabsf(n.z);
    u32 r = (u32)((float)((texel >> 16) & 255u) * s), g = (u32)((float)((texel >> 8) & 255u) * s), b = (u32)((float)(texel & 255u) * s);
    return (r << 16) | (g << 8) | b;
}

// BuildTree.cu:367-499 bmMarchKernel for one ray.  mode 0 = reference semantics (stop at the first leaf that produced a hit);
// mode 1 = visit every leaf the ray's boxes reach and keep the smallest t >= 0 (what "closest hit" would mean without the
// early exit; used only to REPORT how often the reference's answer is the true closest hit).
Hit march_ray(const Scene& S, V3 eye, V3 dir, int mode, u64* counters)
{
    V3 invDir = v3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
    StackNode stack[80];
    stack[0] = { S.wmin, S.wmax, 0, 0, 0 };
    i32 top = 0;
    Hit best = { FLT_MAX, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, 0 };
    u64 nv = 0, nt = 0;
    do
    {
        StackNode st = stack[top--];
        const Node& cur = S.nodes[st.node];
        nv++;
        float boxDist = box_ray(st.mn, st.mx, eye, invDir);
        if (boxDist != FLT_MAX)
        {
            if (cur.left < 0 && cur.right < 0)
            {
                if (cur.leaf >= 0)
                {
                    // already capped to MAX_FACES_PER_BOX (:411)
                    for (u64 k = S.leafFirst[cur.leaf], ke = k + S.leafCount[cur.leaf]; k < ke; k++)
                    {
                        u32 r = S.leafFlat[k];
                        const FaceRec& f = S.faces[r];
                        const Mesh& M = S.meshes[f.mesh];
                        float u, v;
                        float d = tri_intersect(eye, dir, M.pos[f.i0], M.pos[f.i1], M.pos[f.i2], u, v);   // bmFaceRayIntersect :243-251
                        nt++;
                        if (mode == 1 && !(d >= 0.f)) continue;
                        if (d < best.t) { best.t = d; best.u = u; best.v = v; best.mesh = f.mesh; best.tri = f.tri; best.tie = 0; }
                        else if (d == best.t && d != FLT_MAX && !(f.mesh == best.mesh && f.tri == best.tri)) best.tie = 1;
                    }
                    if (mode == 0 && best.t != FLT_MAX) break;            // :427-431
                }
            }
            else
            {
                V3 stMin = st.mn, stMax = st.mx;
                u32 axis = st.axis, naxis = (axis + 1) % 3;
                float s = .5f * (stMax[axis] + stMin[axis]);
                float p = eye[axis] + boxDist * dir[axis];
                if (p < s)
                {
                    if (cur.right >= 0) { float t = stMin[axis]; stMin[axis] = s; stack[++top] = { stMin, stMax, cur.right, 0, naxis }; stMin[axis] = t; }
                    if (cur.left >= 0)  { stMax[axis] = s; stack[++top] = { stMin, stMax, cur.left, 0, naxis }; }
                }
                else
                {
                    if (cur.left >= 0)  { float t = stMax[axis]; stMax[axis] = s; stack[++top] = { stMin, stMax, cur.left, 0, naxis }; stMax[axis] = t; }
                    if (cur.right >= 0) { stMin[axis] = s; stack[++top] = { stMin, stMax, cur.right, 0, naxis }; }
                }
            }
        }
    } while (top >= 0);
    if (counters) { counters[0] += nv; counters[1] += nt; }
    return best;
}

template <class F> void parallel_for(u64 n, int nthreads, F f)
{
    if (nthreads <= 1 || n < 1024) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    u64 chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++)
    {
        u64 a = std::min<u64>(n, t * chunk), b = std::min<u64>(n, a + chunk);
        th.emplace_back([=] { f(a, b, t); });
    }
    for (auto& t : th) t.join();
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
} // namespace

extern "C"
{
int orc_tri_box_overlap(const float* c, const float* h, const float* t9)
{
    V3 t[3] = { v3(t9[0], t9[1], t9[2]), v3(t9[3], t9[4], t9[5]), v3(t9[6], t9[7], t9[8]) };
    return tri_box_overlap(c, h, t);
}
float orc_tri_intersect(const float* o, const float* d, const float* a, const float* b, const float* c, float* u, float* v)
{
    return tri_intersect(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2]), v3(c[0], c[1], c[2]), *u, *v);
}
float orc_box_ray(const float* mn, const float* mx, const float* o, const float* id)
{
    return box_ray(v3(mn[0], mn[1], mn[2]), v3(mx[0], mx[1], mx[2]), v3(o[0], o[1], o[2]), v3(id[0], id[1], id[2]));
}
float orc_box_ray_nozero(const float* mn, const float* mx, const float* o, const float* id)
{
    return box_ray_nozero(v3(mn[0], mn[1], mn[2]), v3(mx[0], mx[1], mx[2]), v3(o[0], o[1], o[2]), v3(id[0], id[1], id[2]));
}
u32 orc_hash(u32 h) { return hash32(h); }
u32 orc_hash3(u32 x, u32 y, u32 z) { return hash3(x, y, z); }
i32 orc_map(float f) { return map_cell(f); }

// Camera.cpp:43-72 Camera::setInitialRays (host loop with accumulating rx / ry)
int orc_initial_rays(u32 width, u32 height, float left, float right, float top, float bottom, float zoom, float* out)
{
    if (width == 0 || height == 0) return 2;
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
            if (std::isnan(d) || d <= 0.f) return 2;
            size_t a = (size_t)y * width + x;
            out[3 * a] = rx * d; out[3 * a + 1] = ry * d; out[3 * a + 2] = zoom * d;
        }
    }
    return 0;
}

// world box / constants default to SceneTree.cpp:43-45 and BuildTree.cuh:14-18 when the pointers are null / values 0
void* orc_scene_create(const float* wmin, const float* wmax, float minLeaf, u32 maxDepth, u32 maxFacesPerLeaf)
{
    Scene* S = new Scene();
    S->wmin = wmin ? v3(wmin[0], wmin[1], wmin[2]) : v3(-30, -30, -30);
    S->wmax = wmax ? v3(wmax[0], wmax[1], wmax[2]) : v3(30, 30, 30);
    S->minLeaf = minLeaf > 0 ? minLeaf : .03f;
    S->maxDepth = maxDepth ? maxDepth : 38;
    S->maxFacesPerLeaf = maxFacesPerLeaf ? maxFacesPerLeaf : 256;
    S->nodes.push_back({ -1, -1, -1, 0 });
    return S;
}
void orc_scene_destroy(void* h) { delete (Scene*)h; }

int orc_scene_add_mesh(void* h, const float* pos, const float* nrm, u32 numVerts, const u32* idx, u32 numIdx)
{
    Scene* S = (Scene*)h;
    if (!pos || !idx || numIdx % 3) return 2;
    Mesh M;
    M.pos.resize(numVerts); memcpy(M.pos.data(), pos, (size_t)numVerts * 12);
    if (nrm) { M.nrm.resize(numVerts); memcpy(M.nrm.data(), nrm, (size_t)numVerts * 12); }
    M.idx.assign(idx, idx + numIdx);
    S->meshes.push_back(std::move(M));
    return 0;
}

int orc_scene_set_mesh_uv(void* h, u32 mesh, const float* uv, u32 comps)
{
    Scene* S = (Scene*)h;
    if (mesh >= S->meshes.size() || !uv || comps < 1 || comps > 4) return 2;
    Mesh& M = S->meshes[mesh];
    M.uv.assign(uv, uv + (size_t)comps * M.pos.size()); M.uvComps = comps;
    return 0;
}

int orc_scene_set_mesh_texture(void* h, u32 mesh, const u32* texels, u32 width, u32 height)
{
    Scene* S = (Scene*)h;
    if (mesh >= S->meshes.size() || !texels || !width || !height) return 2;
    Mesh& M = S->meshes[mesh];
    M.texels.assign(texels, texels + (size_t)width * height); M.texW = width; M.texH = height;
    return 0;
}

// pixels of hit records (u, v, mesh, triangle-in-mesh; tri = 0xFFFFFFFF on a miss) under the material rule above
void orc_shade_hits(void* h, u64 n, const float* u, const float* v, const u32* mesh, const u32* tri, u32* pixels)
{
    Scene* S = (Scene*)h;
    for (u64 i = 0; i < n; i++)
    {
        Hit hit = { 0.f, u[i], v[i], mesh[i], tri[i], 0 };
        pixels[i] = shade_textured(*S, hit);
    }
}

// SceneTree::updateGPUScene (SceneTree.cpp:70-91): reset, then insert every mesh in order. Returns seconds.
double orc_scene_build(void* h)
{
    Scene* S = (Scene*)h;
    double t0 = now();
    S->nodes.clear(); S->nodes.push_back({ -1, -1, -1, 0 });
    S->faces.clear(); S->leafFaces.clear();
    for (u32 m = 0; m < S->meshes.size(); m++)
        for (u32 t = 0; t < S->meshes[m].idx.size() / 3; t++) insert_triangle(*S, m, t);
    flatten_leaf_lists(*S);
    double t1 = now();
    canonical_dump(*S);
    return t1 - t0;
}

// The same tree built with `nthreads` host threads -- for the large-scene parity tests and the CPU baseline of the bench, where the serial
// insertion (0.1-0.3 Mtris/s) would take minutes.  It relies on one property of bmInsertTriangleInTree: WHICH leaves a triangle reaches
// depends only on the triangle and the (fixed) boxes, never on the tree built so far (BuildTree.cu:196-253 tests geometry only; nodes are
// created on demand).  So threads collect (leaf path, mesh, triangle) pairs for disjoint triangle ranges with the very same descent
// (collect_triangle), the pairs are sorted by (path, mesh, triangle) = every leaf's insertion order, and the nodes are created along the
// sorted paths.  Node INDICES differ from the serial build (allocation order), nothing observable does: tests/test_oracle_golden.py
// asserts equal pair lists, branch lists, statistics and traced hits against orc_scene_build.
double orc_scene_build_mt(void* h, int nthreads)
{
    Scene* S = (Scene*)h;
    if (nthreads < 1) nthreads = 1;
    double t0 = now();
    struct Pr { u64 path; u32 mesh, tri; };
    std::vector<u64> meshBase(S->meshes.size() + 1, 0);
    for (size_t m = 0; m < S->meshes.size(); m++) meshBase[m + 1] = meshBase[m] + S->meshes[m].idx.size() / 3;
    const u64 T = meshBase.back();
    std::vector<std::vector<Pr>> per(nthreads);
    u32 leafDepth = 0;
    {
        // depth at which the descent stops: a function of the world box only (every node of a depth has the same box size)
        V3 mn = S->wmin, mx = S->wmax; u32 d = 0, ax = 0;
        for (;;)
        {
            V3 b = mx - mn;
            if (fmin_(b.x, fmin_(b.y, b.z)) < S->minLeaf || d == S->maxDepth - 1) break;
            mx[ax] = .5f * (mn[ax] + mx[ax]); ax = (ax + 1) % 3; d++;
        }
        leafDepth = d;
    }
    parallel_for(T, nthreads, [&](u64 a, u64 b, int tid) {
        std::vector<Pr>& out = per[tid];
        std::vector<StackNode> stack(S->maxDepth + 2);
        std::vector<u64> paths(S->maxDepth + 2);
        size_t m = std::upper_bound(meshBase.begin(), meshBase.end(), a) - meshBase.begin() - 1;
        for (u64 g = a; g < b; g++)
        {
            while (g >= meshBase[m + 1]) m++;
            const Mesh& M = S->meshes[m];
            u32 triIdx = (u32)(g - meshBase[m]);
            V3 v[3] = { M.pos[M.idx[3 * triIdx]], M.pos[M.idx[3 * triIdx + 1]], M.pos[M.idx[3 * triIdx + 2]] };
            // insert_triangle's loop with the node bookkeeping replaced by the path bits (0 = left, 1 = right)
            stack[0] = { S->wmin, S->wmax, 0, 0, 0 }; paths[0] = 0;
            i32 top = 0;
            do
            {
                StackNode st = stack[top]; u64 path = paths[top]; top--;
                V3 bSize = st.mx - st.mn;
                float dMin = fmin_(bSize.x, fmin_(bSize.y, bSize.z));
                if ((dMin < S->minLeaf) || (st.depth == S->maxDepth - 1)) out.push_back({ path, (u32)m, triIdx });
                else
                {
                    u32 ax = st.axis, nAxis = (ax + 1) % 3, ndepth = st.depth + 1;
                    float s = .5f * (st.mn[ax] + st.mx[ax]);
                    V3 lMax = st.mx, rMin = st.mn;
                    lMax[ax] = s; rMin[ax] = s;
                    V3 bc = (lMax + st.mn) * .5f, hs = (lMax - st.mn) * .5f;
                    bool b1 = tri_box_overlap(&bc.x, &hs.x, v) != 0;
                    bc = (st.mx + rMin) * .5f; hs = (st.mx - rMin) * .5f;
                    bool b2 = tri_box_overlap(&bc.x, &hs.x, v) != 0;
                    if (b1) { ++top; stack[top] = { st.mn, lMax, 0, ndepth, nAxis }; paths[top] = path << 1; }
                    if (b2) { ++top; stack[top] = { rMin, st.mx, 0, ndepth, nAxis }; paths[top] = (path << 1) | 1; }
                }
            } while (top >= 0);
        }
    });
    // bucket by the top bits of the path, sort the buckets in parallel: (path, mesh, tri) ascending
    const int BB = leafDepth > 12 ? 12 : (int)leafDepth;
    const u64 NB = 1ull << BB;
    const int sh = (int)leafDepth - BB;
    std::vector<u64> cnt(NB + 1, 0);
    for (auto& v : per) for (auto& p : v) cnt[(p.path >> sh) + 1]++;
    for (u64 i = 0; i < NB; i++) cnt[i + 1] += cnt[i];
    const u64 P = cnt[NB];
    std::vector<Pr> all(P);
    {
        std::vector<u64> cur(cnt.begin(), cnt.end() - 1);
        for (auto& v : per) { for (auto& p : v) all[cur[p.path >> sh]++] = p; std::vector<Pr>().swap(v); }     // threads hold ascending triangle ranges: stable
    }
    parallel_for(NB, nthreads, [&](u64 a, u64 b, int) {
        for (u64 k = a; k < b; k++)
            std::sort(all.begin() + cnt[k], all.begin() + cnt[k + 1], [](const Pr& x, const Pr& y) {
                if (x.path != y.path) return x.path < y.path; if (x.mesh != y.mesh) return x.mesh < y.mesh; return x.tri < y.tri; });
    });
    // nodes along the sorted paths; leaf lists capped like bmTreeNode::insertFace (first maxFacesPerLeaf insertions kept)
    S->nodes.clear(); S->nodes.push_back({ -1, -1, -1, 0 });
    S->faces.clear(); S->leafFaces.clear(); S->leafFlat.clear(); S->leafCount.clear(); S->leafFirst.clear();
    S->faces.reserve(P); S->leafFlat.reserve(P);
    std::vector<i32> pathNode(leafDepth + 1, 0);
    u64 prev = 0; bool have = false;
    for (u64 i = 0; i < P; i++)
    {
        const Pr& p = all[i];
        const Mesh& M = S->meshes[p.mesh];
        if (!have || p.path != prev)
        {
            u32 common = 0;
            if (have) { u64 x = prev ^ p.path; u32 hb = 63 - (u32)__builtin_clzll(x); common = leafDepth - 1 - hb; }   // levels shared with the previous leaf
            for (u32 d = common; d < leafDepth; d++)
            {
                int bit = (int)((p.path >> (leafDepth - 1 - d)) & 1);
                S->nodes.push_back({ -1, -1, -1, 0 });
                i32 nn = (i32)S->nodes.size() - 1;
                if (bit) S->nodes[pathNode[d]].right = nn; else S->nodes[pathNode[d]].left = nn;
                pathNode[d + 1] = nn;
            }
            Node& leaf = S->nodes[pathNode[leafDepth]];
            leaf.leaf = (i32)S->leafCount.size();
            S->leafFirst.push_back(S->leafFlat.size()); S->leafCount.push_back(0);
            prev = p.path; have = true;
        }
        Node& leaf = S->nodes[pathNode[leafDepth]];
        u32 rec = (u32)S->faces.size();
        S->faces.push_back({ M.idx[3 * p.tri], M.idx[3 * p.tri + 1], M.idx[3 * p.tri + 2], p.mesh, p.tri });
        if (leaf.faceInsertIdx++ < S->maxFacesPerLeaf) { S->leafFlat.push_back(rec); S->leafCount.back()++; }
    }
    double t1 = now();
    canonical_dump(*S);
    return t1 - t0;
}

// stats: pairs, non-root nodes, leaves, max leaf depth, min leaf depth, leaves over the face cap, max faces in a leaf, two-child nodes
void orc_scene_stats(void* h, u64* s)
{
    Scene* S = (Scene*)h;
    s[0] = S->pc.size(); s[1] = S->nodes.size() - 1; s[2] = S->nLeaves; s[3] = S->maxLeafDepth; s[4] = S->minLeafDepth;
    s[5] = S->overflowLeaves; s[6] = S->maxLeafFaces; s[7] = S->bdepth.size();
}
void orc_scene_pairs(void* h, u32* code, u32* mesh, u32* tri)
{
    Scene* S = (Scene*)h; size_t n = S->pc.size();
    memcpy(code, S->pc.data(), n * 4); memcpy(mesh, S->pm.data(), n * 4); memcpy(tri, S->pt.data(), n * 4);
}
// nodes of the reference tree that have TWO children, as (depth, path prefix), sorted: the topology a radix tree over the
// distinct leaf codes must reproduce (SURVEY.md section 7.1 "Topology")
void orc_scene_branches(void* h, u32* depth, u32* prefix)
{
    Scene* S = (Scene*)h; size_t n = S->bdepth.size();
    memcpy(depth, S->bdepth.data(), n * 4); memcpy(prefix, S->bprefix.data(), n * 4);
}

// Explicit ray batch: org/dir are n x 3 floats. Outputs may be null. counters[0] += node visits, counters[1] += triangle tests.
double orc_trace_rays(void* h, const float* org, const float* dir, u64 n, int mode, float* t, float* u, float* v, u32* mesh, u32* tri,
                      u32* tie, u32* pixels, u64* counters, int nthreads)
{
    Scene* S = (Scene*)h;
    if (nthreads < 1) nthreads = 1;
    std::vector<u64> cnt((size_t)nthreads * 2, 0);
    double t0 = now();
    parallel_for(n, nthreads, [&](u64 a, u64 b, int tid) {
        for (u64 i = a; i < b; i++)
        {
            Hit hit = march_ray(*S, v3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), v3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), mode, &cnt[2 * tid]);
            if (t) t[i] = hit.t; if (u) u[i] = hit.u; if (v) v[i] = hit.v; if (mesh) mesh[i] = hit.mesh; if (tri) tri[i] = hit.tri;
            if (tie) tie[i] = hit.tie; if (pixels) pixels[i] = shade(*S, hit);
        }
    });
    double t1 = now();
    if (counters) for (int k = 0; k < nthreads; k++) { counters[0] += cnt[2 * k]; counters[1] += cnt[2 * k + 1]; }
    return t1 - t0;
}

// bmMarch (BuildTree.cu:502-520): ray i = orient * initialRays[i] from eye; orient is column-major 3x3 (glm::mat3).
double orc_trace_camera(void* h, const float* initialRays, u64 n, const float* eye3, const float* orient9, int mode, float* t, float* u, float* v,
                        u32* mesh, u32* tri, u32* tie, u32* pixels, u64* counters, int nthreads)
{
    std::vector<float> org(3 * n), dir(3 * n);
    const float* m = orient9;
    for (u64 i = 0; i < n; i++)
    {
        float x = initialRays[3 * i], y = initialRays[3 * i + 1], z = initialRays[3 * i + 2];
        dir[3 * i]     = m[0] * x + m[3] * y + m[6] * z;            // glm mat3 * vec3 (type_mat3x3.inl:424-430)
        dir[3 * i + 1] = m[1] * x + m[4] * y + m[7] * z;
        dir[3 * i + 2] = m[2] * x + m[5] * y + m[8] * z;
        org[3 * i] = eye3[0]; org[3 * i + 1] = eye3[1]; org[3 * i + 2] = eye3[2];
    }
    return orc_trace_rays(h, org.data(), dir.data(), n, mode, t, u, v, mesh, tri, tie, pixels, counters, nthreads);
}
} // extern "C"
