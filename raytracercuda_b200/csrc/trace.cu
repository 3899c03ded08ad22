// trace.cu -- ray / triangle closest-hit traversal of the 64-byte-node radix tree (replaces bmMarchKernel, BuildTree.cu:367-499),
// the pixel shade of BuildTree.cu:486-496, bmClear (RTClear.cu) and the spatial-hash key kernel (Hash.cu:17-64).
//
// BM_TRACE_REFERENCE reproduces the reference's traversal decisions on the contracted tree:
//   * a node is entered only if the ray hits the node's CELL box (the reference's implicit box at that depth), tested with the
//     reference's slab arithmetic (bmBoxRayIntersect); nested cells make the skipped single-child chain boxes redundant;
//   * children are ordered by which side of the split plane the cell-entry point lies (BuildTree.cu:440-481);
//   * a leaf tests its faces in (mesh, triangle) order, keeps the strict minimum d of any sign, at most 256 faces, and the
//     first leaf that produced a hit ends the traversal (BuildTree.cu:408-432);
//   * additionally a child whose tight box (union of its whole triangles) is missed by the ray's LINE is not pushed: no
//     triangle below it can be hit, so results are unchanged and most of the reference's ~97 node visits per ray vanish.
#include "bm_device.cuh"

namespace bm
{
namespace
{
constexpr int TR_THREADS = 128;
constexpr int TR_STACK = 40;     // radix tree depth <= D <= 32

struct HitRec { float t, u, v; u32 tri, mesh, face; };

__device__ __forceinline__ V3 f4xyz(float4 a) { return mk3(a.x, a.y, a.z); }

__device__ __forceinline__ void leaf_faces(const TreeView& tv, u32 first, u32 cnt, V3 o, V3 d, bool forwardOnly, HitRec& best, u32& nt)
{
    const float4* p = tv.ltris + 3 * (size_t)first;
    for (u32 i = 0; i < cnt; i++, p += 3)
    {
        float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        float u, v;
        float dd = tri_intersect(o, d, f4xyz(a), f4xyz(b), f4xyz(c), u, v);
        nt++;
        if (forwardOnly && !(dd >= 0.f)) continue;
        if (dd < best.t)
        {
            best.t = dd; best.u = u; best.v = v;
            best.tri = __float_as_uint(a.w); best.mesh = __float_as_uint(b.w); best.face = __float_as_uint(c.w);
        }
    }
}

__device__ HitRec traverse_reference(const TreeView& tv, V3 o, V3 d, u32& nv, u32& nt)
{
    HitRec best = { FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };
    if (tv.numLeaves == 0) return best;
    const V3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    const Grid& g = tv.grid;
    u32 stack[TR_STACK];
    int top = 0;
    stack[0] = tv.rootRef;
    while (top >= 0)
    {
        u32 ref = stack[top--];
        nv++;
        if (ref & LEAF_BIT)
        {
            uint4 lf = __ldg(tv.leaves + (ref & ~LEAF_BIT));
            u32 ix, iy, iz; decode_path(g.D, lf.x, ix, iy, iz);
            V3 mn, mx; cell_box(g, g.D, ix, iy, iz, mn, mx);
            if (box_ray(mn, mx, o, inv) == FLT_MAX) continue;                     // BuildTree.cu:405-406
            leaf_faces(tv, lf.y, min(lf.z, tv.maxLeafFaces), o, d, false, best, nt);
            if (best.t != FLT_MAX) break;                                          // :427-431
        }
        else
        {
            const float4* np = tv.nodes + 4 * (size_t)ref;
            float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
            u32 c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y), prefix = __float_as_uint(q3.z);
            int plen = (int)__float_as_uint(q3.w);
            u32 ix, iy, iz; decode_path(plen, prefix, ix, iy, iz);
            V3 mn, mx; cell_box(g, plen, ix, iy, iz, mn, mx);
            float boxDist = box_ray(mn, mx, o, inv);
            if (boxDist == FLT_MAX) continue;
            int axis = plen % 3;
            u32 ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
            int sh = g.k[axis] - axis_bits(plen, axis);                            // >= 1: the node splits on this axis
            float s = __ldg(g.tab[axis] + (((ia << 1) | 1u) << (sh - 1)));         // == .5f*(max+min) of the cell, :443
            float p = get(o, axis) + boxDist * get(d, axis);                       // :444
            bool leftFirst = p < s;
            bool h0 = line_hits_box(mk3(q0.x, q0.y, q0.z), mk3(q0.w, q1.x, q1.y), o, inv);
            bool h1 = line_hits_box(mk3(q1.z, q1.w, q2.x), mk3(q2.y, q2.z, q2.w), o, inv);
            u32 nearRef = leftFirst ? c0 : c1, farRef = leftFirst ? c1 : c0;
            bool nearHit = leftFirst ? h0 : h1, farHit = leftFirst ? h1 : h0;
            if (farHit) stack[++top] = farRef;
            if (nearHit) stack[++top] = nearRef;
        }
    }
    return best;
}

// tight-box slab for the exhaustive mode: forward interval [0, tmax]
__device__ __forceinline__ bool box_interval(V3 bMin, V3 bMax, V3 o, V3 inv, float tmaxRay, float& tnear)
{
    V3 t0 = (bMin - o) * inv, t1 = (bMax - o) * inv;
    float tn = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fmaxf(fminf(t0.z, t1.z), 0.f));
    float tf = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fminf(fmaxf(t0.z, t1.z), tmaxRay));
    tnear = tn;
    return tn <= tf;
}

__device__ HitRec traverse_closest(const TreeView& tv, V3 o, V3 d, u32& nv, u32& nt)
{
    HitRec best = { FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };
    if (tv.numLeaves == 0) return best;
    const V3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    u32 stack[TR_STACK]; float stackT[TR_STACK];
    int top = 0;
    stack[0] = tv.rootRef; stackT[0] = 0.f;
    while (top >= 0)
    {
        u32 ref = stack[top]; float tn = stackT[top]; --top;
        if (tn > best.t) continue;
        nv++;
        if (ref & LEAF_BIT)
        {
            uint4 lf = __ldg(tv.leaves + (ref & ~LEAF_BIT));
            leaf_faces(tv, lf.y, lf.z, o, d, true, best, nt);
        }
        else
        {
            const float4* np = tv.nodes + 4 * (size_t)ref;
            float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
            u32 c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y);
            float t0, t1;
            bool h0 = box_interval(mk3(q0.x, q0.y, q0.z), mk3(q0.w, q1.x, q1.y), o, inv, best.t, t0);
            bool h1 = box_interval(mk3(q1.z, q1.w, q2.x), mk3(q2.y, q2.z, q2.w), o, inv, best.t, t1);
            if (h0 && h1)
            {
                bool firstIs0 = t0 <= t1;
                ++top; stack[top] = firstIs0 ? c1 : c0; stackT[top] = firstIs0 ? t1 : t0;
                ++top; stack[top] = firstIs0 ? c0 : c1; stackT[top] = firstIs0 ? t0 : t1;
            }
            else if (h0) { ++top; stack[top] = c0; stackT[top] = t0; }
            else if (h1) { ++top; stack[top] = c1; stackT[top] = t1; }
        }
    }
    return best;
}

// bmFaceInterpolate<vec3>(face, u, v, meshData, 1) + normalize + pack (CudaComon.cuh:253-266, BuildTree.cu:486-496).
// A mesh without normals (null slot 1; undefined behaviour in the reference) shades with the geometric normal.
__device__ u32 shade(const HitRec& h, const bmStaticMeshData* __restrict__ meshData)
{
    if (h.tri == 0xFFFFFFFFu) return 255u << 8;
    const bmStaticMeshData* m = meshData + h.mesh;
    const u32* idx = m->m_indices + 3 * (size_t)h.face;
    u32 i0 = idx[0], i1 = idx[1], i2 = idx[2];
    const float* nrm = m->m_vertexData[1];
    V3 n;
    if (nrm)
    {
        V3 n0 = mk3(nrm[3 * i0], nrm[3 * i0 + 1], nrm[3 * i0 + 2]);
        V3 n1 = mk3(nrm[3 * i1], nrm[3 * i1 + 1], nrm[3 * i1 + 2]);
        V3 n2 = mk3(nrm[3 * i2], nrm[3 * i2 + 1], nrm[3 * i2 + 2]);
        float w = 1 - (h.u + h.v);
        n = n0 * w + n1 * h.u + n2 * h.v;
    }
    else
    {
        const float* pos = m->m_vertexData[0];
        V3 p0 = mk3(pos[3 * i0], pos[3 * i0 + 1], pos[3 * i0 + 2]);
        V3 p1 = mk3(pos[3 * i1], pos[3 * i1 + 1], pos[3 * i1 + 2]);
        V3 p2 = mk3(pos[3 * i2], pos[3 * i2 + 1], pos[3 * i2 + 2]);
        n = cross3(p1 - p0, p2 - p0);
    }
    n = n * (1.f / sqrtf(dot3(n, n)));          // glm::normalize = v * inversesqrt(dot(v, v)), inversesqrt = 1/sqrt
    return (u32)fabsf(n.z * 255) << 16;
}

template <bool CAMERA>
__global__ void __launch_bounds__(TR_THREADS) trace_kernel(TreeView tv, const bmRay* __restrict__ rays, CameraIn cam, u64 n, int mode,
                                                           bmHit* __restrict__ hits, u32* __restrict__ pixels,
                                                           const bmStaticMeshData* __restrict__ meshData, u64* counters)
{
    u32 nv = 0, nt = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    {
        V3 o, d;
        if (CAMERA)
        {
            const float* r = cam.initialRays + 3 * i;
            float x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
            const float* m = cam.orient;                                      // glm mat3 * vec3 (type_mat3x3.inl:424-430)
            d = mk3(m[0] * x + m[3] * y + m[6] * z, m[1] * x + m[4] * y + m[7] * z, m[2] * x + m[5] * y + m[8] * z);
            o = mk3(cam.eye[0], cam.eye[1], cam.eye[2]);
        }
        else
        {
            const float* r = reinterpret_cast<const float*>(rays + i);
            o = mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2));
            d = mk3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
        }
        HitRec h = (mode == BM_TRACE_REFERENCE) ? traverse_reference(tv, o, d, nv, nt) : traverse_closest(tv, o, d, nv, nt);
        if (hits) { bmHit out; out.t = h.t; out.u = h.u; out.v = h.v; out.tri = h.tri; reinterpret_cast<float4*>(hits)[i] = *reinterpret_cast<float4*>(&out); }
        if (pixels) pixels[i] = shade(h, meshData);
    }
    if (counters)
    {
        u64 a = nv, b = nt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xFFFFFFFFu, a, o); b += __shfl_down_sync(0xFFFFFFFFu, b, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd((unsigned long long*)&counters[0], (unsigned long long)a); atomicAdd((unsigned long long*)&counters[1], (unsigned long long)b); }
    }
}

__global__ void clear_kernel(u32* __restrict__ buf, size_t n, u32 v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}

__global__ void hash_kernel(const float* __restrict__ pts, u64 n, u32* __restrict__ hash, i32* __restrict__ cells)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i32 cx = map_cell(pts[3 * i]), cy = map_cell(pts[3 * i + 1]), cz = map_cell(pts[3 * i + 2]);
    if (cells) { cells[3 * i] = cx; cells[3 * i + 1] = cy; cells[3 * i + 2] = cz; }
    if (hash) hash[i] = hash3((u32)cx, (u32)cy, (u32)cz);
}

unsigned trace_grid(u64 n)
{
    u64 blocks = (n + TR_THREADS - 1) / TR_THREADS;
    u64 cap = (u64)sm_count() * 16;
    return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}
} // namespace

int trace_rays(const TreeView& tv, const bmRay* d_rays, u64 n, int mode, bmHit* d_hits, u64* d_counters, cudaStream_t s)
{
    if (!n) return BM_OK;
    CameraIn cam = {};
    trace_kernel<false><<<trace_grid(n), TR_THREADS, 0, s>>>(tv, d_rays, cam, n, mode, d_hits, nullptr, nullptr, d_counters);
    BM_LAUNCHED();
    return BM_OK;
}

int trace_camera(const TreeView& tv, const CameraIn& cam, u64 n, int mode, bmHit* d_hits, u32* d_pixels, const bmStaticMeshData* d_meshData,
                 u64* d_counters, cudaStream_t s)
{
    if (!n) return BM_OK;
    trace_kernel<true><<<trace_grid(n), TR_THREADS, 0, s>>>(tv, nullptr, cam, n, mode, d_hits, d_pixels, d_meshData, d_counters);
    BM_LAUNCHED();
    return BM_OK;
}

int clear_u32(u32* d_buf, size_t n, u32 value, cudaStream_t s)
{
    if (!n) return BM_OK;
    unsigned blocks = div_up(n, 256);
    unsigned cap = (unsigned)sm_count() * 8;
    clear_kernel<<<blocks < cap ? blocks : cap, 256, 0, s>>>(d_buf, n, value);
    BM_LAUNCHED();
    return BM_OK;
}

int hash_keys(const float* d_points3, u64 n, u32* d_hash, i32* d_cells3, cudaStream_t s)
{
    if (!n) return BM_OK;
    hash_kernel<<<div_up(n, 256), 256, 0, s>>>(d_points3, n, d_hash, d_cells3);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
