// trace.cu -- ray / triangle closest-hit traversal of the 128-byte-node radix tree (replaces bmMarchKernel, BuildTree.cu:367-499),
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
#include <cstdlib>
#include <cstring>
#include "bm_device.cuh"
#include "bm_wide.cuh"

namespace bm
{
namespace
{
// Conservative fused pre-test in front of the exact triangle test (tri_reject, bm_math.cuh): OFF.  It catches 99.9 % of ordinary misses for
// a third of the exact test's instructions (tests/test_hostcheck.py), but the face loops run ~10 lanes at a time and one of them passes
// the pre-test in ~85 % of the trips, so the warp executes both tests: c2 step 0.850 -> 0.884 ms, c4 frame 52.8 -> 56.1 ms with it
// (profiles/r2_SUMMARY.md).  Kept as a compile-time variant (-DBM_TRI_PRETEST=1) with its CPU property test.
#ifndef BM_TRI_PRETEST
#define BM_TRI_PRETEST 0
#endif
constexpr int TR_THREADS = 128;
constexpr int TR_STACK = 52;     // binary walk: depth <= D <= 32; wide walk: a visit pops one entry and pushes up to four, <= 16 wide levels
#ifndef BM_COOP_MAX_LANES
#define BM_COOP_MAX_LANES 8
#endif
#ifndef BM_COOP_MIN_FACES
#define BM_COOP_MIN_FACES 8
#endif
constexpr int COOP_MAX_LANES = BM_COOP_MAX_LANES;   // at most this many pending leaves in a warp -> process the sizeable ones warp-cooperatively
constexpr int COOP_MIN_FACES = BM_COOP_MIN_FACES;   // ... where sizeable = more than this many faces
constexpr int TR_QUEUE_MAX = 8;      // leaves a lane of a ray batch may postpone (queueDepth <= this)
constexpr u32 REF_EMPTY = 0x7FFFFFFFu;   // "nothing left to visit" (never a valid internal index: N < 2^31 - 1)

struct HitRec { float t, u, v; u32 tri, mesh, face; };

__device__ __forceinline__ V3 f4xyz(float4 a) { return mk3(a.x, a.y, a.z); }

// Conservative LINE/box cull on the padded tight boxes, one FMA per slab plane: t = b*ic - o*ic with o*ic hoisted per ray.
// ic is 1/d clamped to +-1e30 (cull_inv), so an axis-parallel ray gives +-huge finite values with the right signs instead of the
// inf - inf = NaN an unclamped reciprocal would produce (a NaN on one side of a slab would silently flip the test).  Rounding
// differs from (b-o)/d by at most |o|*2^-24 in space, far inside the boxes' pad (world extent * 2^-20).  Never decides a result:
// it only skips subtrees / face groups that cannot contain a hit.
__device__ __forceinline__ float cull_inv(float inv) { return (inv == inv) ? fminf(fmaxf(inv, -1e30f), 1e30f) : 1e30f; }
__device__ __forceinline__ bool line_hits_box_fma(V3 bMin, V3 bMax, V3 inv, V3 oi)
{
    float x0 = __fmaf_rn(bMin.x, inv.x, -oi.x), x1 = __fmaf_rn(bMax.x, inv.x, -oi.x);
    float y0 = __fmaf_rn(bMin.y, inv.y, -oi.y), y1 = __fmaf_rn(bMax.y, inv.y, -oi.y);
    float z0 = __fmaf_rn(bMin.z, inv.z, -oi.z), z1 = __fmaf_rn(bMax.z, inv.z, -oi.z);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return !(tf < tn);
}

// the same slabs for the CLOSEST mode's leaf-interior culling: the box must overlap the ray's [0, tbest] (a hit inside the unpadded box
// has a t that every padded slab interval contains, so tf >= t >= 0 and tn <= t <= tbest: conservative like the line test)
__device__ __forceinline__ bool seg_hits_box_fma(V3 bMin, V3 bMax, V3 inv, V3 oi, float tbest)
{
    float x0 = __fmaf_rn(bMin.x, inv.x, -oi.x), x1 = __fmaf_rn(bMax.x, inv.x, -oi.x);
    float y0 = __fmaf_rn(bMin.y, inv.y, -oi.y), y1 = __fmaf_rn(bMax.y, inv.y, -oi.y);
    float z0 = __fmaf_rn(bMin.z, inv.z, -oi.z), z1 = __fmaf_rn(bMax.z, inv.z, -oi.z);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return !(tf < tn) && !(tf < 0.f) && !(tn > tbest);
}
template <int MODE>
__device__ __forceinline__ bool cluster_hit(V3 bMin, V3 bMax, V3 inv, V3 oi, float tbest)
{
    return MODE == BM_TRACE_REFERENCE ? line_hits_box_fma(bMin, bMax, inv, oi) : seg_hits_box_fma(bMin, bMax, inv, oi, tbest);
}

// ---- one step of each traversal flavour, shared by every kernel variant -------------------------------------------------
// Internal node in REFERENCE mode: cell test with the reference's slab arithmetic, child order by the split plane
// (BuildTree.cu:398-481), children whose tight box the ray's line misses are not pushed.
__device__ __forceinline__ void ref_node(const TreeView& tv, u32 ref, V3 o, V3 d, V3 inv, V3 ic, V3 oi, u32* stack, int& top)
{
    const float4* np = tv.nodes + NODE_F4 * (size_t)ref;
    // all six quads at once: the node is one 128-byte L1 line, and a second dependent round trip would double the latency
    float4 q4 = __ldg(np + 4), q5 = __ldg(np + 5), q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
    float boxDist = box_ray(mk3(q4.x, q4.y, q4.z), mk3(q5.x, q5.y, q5.z), o, inv);     // the node's cell, :398
    if (boxDist == FLT_MAX) return;
    u32 c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y);
    int axis = (int)__float_as_uint(q5.w);
    float p = get(o, axis) + boxDist * get(d, axis);                                   // :444
    bool leftFirst = p < q4.w;                                                          // s = .5f*(max+min) of the cell, :443
    bool h0 = line_hits_box_fma(mk3(q0.x, q0.y, q0.z), mk3(q0.w, q1.x, q1.y), ic, oi);
    bool h1 = line_hits_box_fma(mk3(q1.z, q1.w, q2.x), mk3(q2.y, q2.z, q2.w), ic, oi);
    u32 nearRef = leftFirst ? c0 : c1, farRef = leftFirst ? c1 : c0;
    bool nearHit = leftFirst ? h0 : h1, farHit = leftFirst ? h1 : h0;
    if (farHit)
    {
        // Does the forward ray reach the FAR half of this cell at all?  Every leaf cell below the far child lies inside that half, and the
        // reference tests a leaf only if the ray enters its cell (:398-406), so a far half the ray provably misses cannot contribute: the
        // child would be rejected by its own cell test on arrival, and is not pushed.  Cheap slabs (one FMA per plane; five planes are the
        // cell's, the sixth is the split plane) with a margin that covers their rounding, including the cancellation when |o / d| is large:
        // the decision is only ever "provably missed".
        const float sx0 = __fmaf_rn(q4.x, ic.x, -oi.x), sx1 = __fmaf_rn(q5.x, ic.x, -oi.x);
        const float sy0 = __fmaf_rn(q4.y, ic.y, -oi.y), sy1 = __fmaf_rn(q5.y, ic.y, -oi.y);
        const float sz0 = __fmaf_rn(q4.z, ic.z, -oi.z), sz1 = __fmaf_rn(q5.z, ic.z, -oi.z);
        const float ts = __fmaf_rn(q4.w, get(ic, axis), -get(oi, axis));
        // far = upper half when the near one is the lower: its low plane on `axis` is the split plane
        const float ax0 = axis == 0 ? (leftFirst ? ts : sx0) : sx0, ax1 = axis == 0 ? (leftFirst ? sx1 : ts) : sx1;
        const float ay0 = axis == 1 ? (leftFirst ? ts : sy0) : sy0, ay1 = axis == 1 ? (leftFirst ? sy1 : ts) : sy1;
        const float az0 = axis == 2 ? (leftFirst ? ts : sz0) : sz0, az1 = axis == 2 ? (leftFirst ? sz1 : ts) : sz1;
        const float tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fminf(az0, az1));
        const float tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
        const float margin = 1e-5f * (fabsf(tn) + fabsf(tf) + fabsf(oi.x) + fabsf(oi.y) + fabsf(oi.z)) + 1e-30f;
        if (!(tf + margin < fmaxf(tn, 0.f))) stack[++top] = farRef;
    }
    if (nearHit) stack[++top] = nearRef;
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

// Internal node in CLOSEST mode: both tight boxes against [0, best.t], near child on top of the stack.  A popped subtree
// whose box lies beyond the hit found meanwhile is rejected again when its own node is visited (its children fail).
//
// The stored child boxes are unions of WHOLE triangles (what the reference mode needs: it accepts a hit wherever the line meets a triangle
// of a leaf whose cell the ray enters).  A closest hit with t >= 0 at point P is always found through the leaf whose CELL holds P -- every
// cell a triangle overlaps lists it -- so here each child box is first cut down to the child's half of the node's cell (grown by cellPad
// against rounding): a triangle spanning thousands of cells (f16's wings) no longer drags every ray near it through all of its leaves.
__device__ __forceinline__ void closest_node(const TreeView& tv, u32 ref, V3 o, V3 inv, float tbest, u32* stack, int& top)
{
    const float4* np = tv.nodes + NODE_F4 * (size_t)ref;
    float4 q4 = __ldg(np + 4), q5 = __ldg(np + 5), q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
    u32 c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y);
    const float pad = tv.cellPad;
    const int axis = (int)__float_as_uint(q5.w);
    const V3 cmn = mk3(q4.x - pad, q4.y - pad, q4.z - pad), cmx = mk3(q5.x + pad, q5.y + pad, q5.z + pad);
    const float sLo = q4.w - pad, sHi = q4.w + pad;                      // split plane: upper bound of child 0's half, lower bound of child 1's
    const V3 hi0 = mk3(axis == 0 ? sHi : cmx.x, axis == 1 ? sHi : cmx.y, axis == 2 ? sHi : cmx.z);
    const V3 lo1 = mk3(axis == 0 ? sLo : cmn.x, axis == 1 ? sLo : cmn.y, axis == 2 ? sLo : cmn.z);
    float t0, t1;
    bool h0 = box_interval(mk3(fmaxf(q0.x, cmn.x), fmaxf(q0.y, cmn.y), fmaxf(q0.z, cmn.z)), mk3(fminf(q0.w, hi0.x), fminf(q1.x, hi0.y), fminf(q1.y, hi0.z)), o, inv, tbest, t0);
    bool h1 = box_interval(mk3(fmaxf(q1.z, lo1.x), fmaxf(q1.w, lo1.y), fmaxf(q2.x, lo1.z)), mk3(fminf(q2.y, cmx.x), fminf(q2.z, cmx.y), fminf(q2.w, cmx.z)), o, inv, tbest, t1);
    bool firstIs0 = t0 <= t1;
    u32 nearRef = firstIs0 ? c0 : c1, farRef = firstIs0 ? c1 : c0;
    bool nearHit = firstIs0 ? h0 : h1, farHit = firstIs0 ? h1 : h0;
    if (farHit) stack[++top] = farRef;
    if (nearHit) stack[++top] = nearRef;
}

// ---- the 64-byte wide node: one line in, up to four references out (wide_step, bm_wide.cuh) ---------------------------------------
template <int MODE>
__device__ __forceinline__ void wide_node(const TreeView& tv, u32 ref, V3 o, V3 d, V3 inv, V3 ic, V3 oi, V3 fm, float tbest, u32* stack, int& top)
{
    const uint4* np = tv.wnodes + WNODE_U4 * (size_t)ref;
    const uint4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2), e = __ldg(np + 3);
    u32 push[4];
    const int n = wide_step<MODE>(tv.grid, WQ{ a.x, a.y, a.z, a.w }, WQ{ b.x, b.y, b.z, b.w }, WQ{ c.x, c.y, c.z, c.w }, WQ{ e.x, e.y, e.z, e.w }, o, d, inv, ic, oi, fm,
                                  tbest, push);
#pragma unroll
    for (int k = 0; k < 4; k++) if (k < n) stack[++top] = push[k];
}

// One face against the ray.  The reference keeps the strict minimum while walking the leaf's list in insertion order, i.e. on
// equal d the face inserted first = the smaller (mesh, triangle) id wins (BuildTree.cu:413-424); the pairs of a leaf are stored
// in sub-key order here, so the tie is broken on the id explicitly.
template <int MODE>
__device__ __forceinline__ void test_face(float4 a, float4 b, float4 c, V3 o, V3 d, HitRec& best)
{
    u32 meshBits = __float_as_uint(b.w);
    if (MODE == BM_TRACE_REFERENCE && (meshBits & EXCLUDED_BIT)) return;       // beyond the 256-face cap of its leaf
    float u, v;
#if BM_TRI_PRETEST
    if (tri_reject(o, d, f4xyz(a), f4xyz(b), f4xyz(c))) return;       // provably no hit (bm_math.cuh): the exact test and its division are skipped
#endif
    float dd = tri_intersect(o, d, f4xyz(a), f4xyz(b), f4xyz(c), u, v);
    if (MODE != BM_TRACE_REFERENCE && !(dd >= 0.f)) return;
    u32 tri = __float_as_uint(a.w);
    if (dd < best.t || (dd == best.t && dd != FLT_MAX && tri < best.tri))
    {
        best.t = dd; best.u = u; best.v = v;
        best.tri = tri; best.mesh = meshBits & ~EXCLUDED_BIT; best.face = __float_as_uint(c.w);
    }
}

// bmFaceInterpolate<vec3>(face, u, v, meshData, 1) + normalize (CudaComon.cuh:253-266, BuildTree.cu:486-492): the z component of
// the shading normal.  A mesh without normals (null slot 1; undefined behaviour in the reference) uses the geometric normal.
__device__ float shade_nz(u32 mesh, u32 face, float u, float v, const bmStaticMeshData* __restrict__ meshData, u32& i0, u32& i1, u32& i2)
{
    const bmStaticMeshData* m = meshData + mesh;
    const u32* idx = m->m_indices + 3 * (size_t)face;
    i0 = idx[0]; i1 = idx[1]; i2 = idx[2];
    const float* nrm = m->m_vertexData[1];
    V3 n;
    if (nrm)
    {
        V3 n0 = mk3(nrm[3 * i0], nrm[3 * i0 + 1], nrm[3 * i0 + 2]);
        V3 n1 = mk3(nrm[3 * i1], nrm[3 * i1 + 1], nrm[3 * i1 + 2]);
        V3 n2 = mk3(nrm[3 * i2], nrm[3 * i2 + 1], nrm[3 * i2 + 2]);
        float w = 1 - (u + v);
        n = n0 * w + n1 * u + n2 * v;
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
    return n.z;
}

// the reference's pixel (BuildTree.cu:486-496)
__device__ u32 shade(const HitRec& h, const bmStaticMeshData* __restrict__ meshData)
{
    if (h.tri == 0xFFFFFFFFu) return 255u << 8;
    u32 i0, i1, i2;
    return (u32)fabsf(shade_nz(h.mesh, h.face, h.u, h.v, meshData, i0, i1, i2) * 255) << 16;
}

// Deferred shading of hit records with the material path the reference declares but never implements (SharedTypes.h:20-23
// bmMaterial{int texture}, Beam.h:21 VERTEX_DATA_UV1, Content/f16.mtl map_Kd): the behaviour is defined here.
//   miss                                  -> 255 << 8                                    (the reference's miss colour)
//   hit, mesh without texture or UV slot  -> (u32)|n.z * 255| << 16                      (the reference's pixel, bit for bit)
//   hit, textured mesh                    -> uv = bmFaceInterpolate(slot 2); texel = nearest, repeat wrap, v up (row 0 = top of the
//                                            image); channel c -> (u32)(c * |n.z|), packed 0x00RRGGBB like the reference's red
__global__ void shade_hits_kernel(const bmHit* __restrict__ hits, u64 n, const float4* __restrict__ soup, const bmStaticMeshData* __restrict__ meshData,
                                  const bmMaterial* __restrict__ materials, const bmTexture* __restrict__ textures, u32 numTextures, u32* __restrict__ pixels)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 h = reinterpret_cast<const float4*>(hits)[i];
    u32 tri = __float_as_uint(h.w);
    if (tri == 0xFFFFFFFFu) { pixels[i] = 255u << 8; return; }
    u32 mesh = __float_as_uint(soup[3 * (size_t)tri + 1].w), face = __float_as_uint(soup[3 * (size_t)tri + 2].w);
    u32 i0, i1, i2;
    float nz = shade_nz(mesh, face, h.y, h.z, meshData, i0, i1, i2);
    int tex = materials ? materials[mesh].texture : -1;
    const bmStaticMeshData* m = meshData + mesh;
    const float* uv = m->m_vertexData[2];
    u32 comps = m->m_vertexDataSizes[2];
    if (tex < 0 || (u32)tex >= numTextures || !uv || comps < 2) { pixels[i] = (u32)fabsf(nz * 255) << 16; return; }
    bmTexture t = textures[tex];
    float w = 1 - (h.y + h.z);
    float tu = uv[comps * (size_t)i0] * w + uv[comps * (size_t)i1] * h.y + uv[comps * (size_t)i2] * h.z;
    float tv = uv[comps * (size_t)i0 + 1] * w + uv[comps * (size_t)i1 + 1] * h.y + uv[comps * (size_t)i2 + 1] * h.z;
    tu = tu - floorf(tu); tv = 1.f - (tv - floorf(tv));
    u32 x = (u32)(tu * (float)t.width), y = (u32)(tv * (float)t.height);
    x = x < t.width ? x : t.width - 1; y = y < t.height ? y : t.height - 1;
    u32 texel = t.texels[(size_t)y * t.width + x];
    float s = fabsf(nz);
    u32 r = (u32)((float)((texel >> 16) & 255u) * s), g = (u32)((float)((texel >> 8) & 255u) * s), b = (u32)((float)(texel & 255u) * s);
    pixels[i] = (r << 16) | (g << 8) | b;
}

// monotone map float -> u32 (any sign, -0 folded onto +0) so that (d, triangle id) pairs can be min-reduced as one 64-bit key
__device__ __forceinline__ u32 orderable(float f)
{
    u32 u = __float_as_uint(f == 0.f ? 0.f : f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- bulk-asynchronous staging of a leaf's run (BM_TMA_LEAF=1 builds) --------------------------------------------------------------
// A leaf's triangle records are ONE contiguous run of 48-byte records and its cluster boxes one run of 32-byte records: the shape the
// bulk copy engine wants.  The warp-cooperative leaf stages up to 32 groups (128 faces, 6 KB) and their boxes (1 KB) into the warp's
// shared-memory buffer with two cp.async.bulk transfers completing on one mbarrier, then tests boxes and faces out of shared memory:
// one asynchronous round trip instead of the two dependent rounds of per-lane loads (boxes, then the faces of the groups that pass).
#ifndef BM_TMA_LEAF
#define BM_TMA_LEAF 0
#endif
constexpr int TMA_GROUPS = 32;                                            // groups staged per step
constexpr int TMA_TRI_BYTES = TMA_GROUPS * CLUSTER * 48, TMA_BOX_BYTES = TMA_GROUPS * 32;
constexpr int TMA_WARP_BYTES = TMA_TRI_BYTES + TMA_BOX_BYTES;            // 7,168 bytes per warp
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// coop_leaf with the run staged through shared memory (same tests, same winner rule)
template <int MODE, bool COUNT>
__device__ __forceinline__ void coop_leaf_staged(const TreeView& tv, unsigned lane, int owner, V3 o, V3 d, V3 ic, V3 oi, u32 first, u32 cnt, u32 cfirst,
                                                 HitRec& best, u32& nt, u32& nc, unsigned char* stage, unsigned long long* bar, u32& phase)
{
    const unsigned FULL = 0xFFFFFFFFu;
    V3 ro = mk3(__shfl_sync(FULL, o.x, owner), __shfl_sync(FULL, o.y, owner), __shfl_sync(FULL, o.z, owner));
    V3 rd = mk3(__shfl_sync(FULL, d.x, owner), __shfl_sync(FULL, d.y, owner), __shfl_sync(FULL, d.z, owner));
    V3 ric = mk3(__shfl_sync(FULL, ic.x, owner), __shfl_sync(FULL, ic.y, owner), __shfl_sync(FULL, ic.z, owner));
    V3 roi = mk3(__shfl_sync(FULL, oi.x, owner), __shfl_sync(FULL, oi.y, owner), __shfl_sync(FULL, oi.z, owner));
    u32 lfirst = __shfl_sync(FULL, first, owner), lcnt = __shfl_sync(FULL, cnt, owner), lcf = __shfl_sync(FULL, cfirst, owner);
    const float rbest = __shfl_sync(FULL, best.t, owner);
    HitRec mine = { FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };
    const u32 ngroups = (lcnt + CLUSTER - 1) / CLUSTER;
    u32 tests = 0, boxes = 0;
    const float4* sTri = reinterpret_cast<const float4*>(stage);
    const float4* sBox = reinterpret_cast<const float4*>(stage + TMA_TRI_BYTES);
    for (u32 gbase = 0; gbase < ngroups; gbase += TMA_GROUPS)
    {
        const u32 gHere = min(ngroups - gbase, (u32)TMA_GROUPS);
        const u32 f0 = gbase * CLUSTER, fHere = min(lcnt - f0, (u32)(TMA_GROUPS * CLUSTER));
        const bool needBoxes = ngroups > 1;
        if (lane == 0)
        {
            // the buffer was last read through the generic proxy: order those reads before the async-proxy writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, fHere * 48u + (needBoxes ? gHere * 32u : 0u));
            bulk_g2s(stage, tv.ltris + 3 * (size_t)(lfirst + f0), fHere * 48u, bar);
            if (needBoxes) bulk_g2s(stage + TMA_TRI_BYTES, tv.clusters + CLUSTER_F4 * (size_t)(lcf + gbase), gHere * 32u, bar);
        }
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1u;
        bool pass = false;
        if (lane < gHere)
        {
            if (!needBoxes) pass = true;
            else
            {
                float4 b0 = sBox[CLUSTER_F4 * lane], b1 = sBox[CLUSTER_F4 * lane + 1];
                boxes++;
                pass = cluster_hit<MODE>(mk3(b0.x, b0.y, b0.z), mk3(b0.w, b1.x, b1.y), ric, roi, rbest);
            }
        }
        unsigned bits = __ballot_sync(FULL, pass);
        u32 npass = __popc(bits);
        for (u32 k0 = 0; k0 < npass; k0 += 32 / CLUSTER)
        {
            u32 k = k0 + lane / CLUSTER;
            if (k < npass)
            {
                u32 grp = __fns(bits, 0, (int)k + 1);
                u32 i = grp * CLUSTER + lane % CLUSTER;
                if (i < fHere)
                {
                    const float4* p = sTri + 3 * i;
                    float4 a = p[0], b = p[1], c = p[2];
                    tests++;
                    test_face<MODE>(a, b, c, ro, rd, mine);
                }
            }
        }
        __syncwarp();
    }
    u32 khi = orderable(mine.t), klo = mine.tri;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
    {
        u32 ohi = __shfl_xor_sync(FULL, khi, m), olo = __shfl_xor_sync(FULL, klo, m);
        if (ohi < khi || (ohi == khi && olo < klo)) { khi = ohi; klo = olo; }
    }
    unsigned who = __ballot_sync(FULL, orderable(mine.t) == khi && mine.tri == klo);
    int w = __ffs(who) - 1;
    float wt = __shfl_sync(FULL, mine.t, w), wu = __shfl_sync(FULL, mine.u, w), wv = __shfl_sync(FULL, mine.v, w);
    u32 wtri = __shfl_sync(FULL, mine.tri, w), wmesh = __shfl_sync(FULL, mine.mesh, w), wface = __shfl_sync(FULL, mine.face, w);
    if (COUNT) { tests = __reduce_add_sync(FULL, tests); boxes = __reduce_add_sync(FULL, boxes); }
    if ((int)lane == owner)
    {
        if (COUNT) { nt += tests; nc += boxes; }
        if (wt < best.t || (wt == best.t && wt != FLT_MAX && wtri < best.tri))
        {
            best.t = wt; best.u = wu; best.v = wv; best.tri = wtri; best.mesh = wmesh; best.face = wface;
        }
    }
}

// Warp-cooperative leaf: all 32 lanes work on ONE ray's leaf -- 32 cluster boxes per step, then 32 faces per step (8 passing groups
// of 4) -- and a 64-bit min-reduction picks the winner under the reference's rule (strict minimum d of any sign, smaller id on a
// tie).  Used when only a few lanes of the warp have a leaf pending (the tail of a launch is a handful of long, grazing rays whose
// work is ~90 % cluster boxes and faces): it turns hundreds of dependent steps of one lane into a few warp-wide ones.
template <int MODE, bool COUNT>
__device__ __forceinline__ void coop_leaf(const TreeView& tv, unsigned lane, int owner, V3 o, V3 d, V3 ic, V3 oi, u32 first, u32 cnt, u32 cfirst,
                                          HitRec& best, u32& nt, u32& nc)
{
    const unsigned FULL = 0xFFFFFFFFu;
    // the owner's ray and leaf, broadcast
    V3 ro = mk3(__shfl_sync(FULL, o.x, owner), __shfl_sync(FULL, o.y, owner), __shfl_sync(FULL, o.z, owner));
    V3 rd = mk3(__shfl_sync(FULL, d.x, owner), __shfl_sync(FULL, d.y, owner), __shfl_sync(FULL, d.z, owner));
    V3 ric = mk3(__shfl_sync(FULL, ic.x, owner), __shfl_sync(FULL, ic.y, owner), __shfl_sync(FULL, ic.z, owner));
    V3 roi = mk3(__shfl_sync(FULL, oi.x, owner), __shfl_sync(FULL, oi.y, owner), __shfl_sync(FULL, oi.z, owner));
    u32 lfirst = __shfl_sync(FULL, first, owner), lcnt = __shfl_sync(FULL, cnt, owner), lcf = __shfl_sync(FULL, cfirst, owner);
    const float rbest = __shfl_sync(FULL, best.t, owner);
    HitRec mine = { FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };
    u32 ngroups = (lcnt + CLUSTER - 1) / CLUSTER;
    u32 tests = 0, boxes = 0;
    for (u32 gbase = 0; gbase < ngroups; gbase += 32)
    {
        // 32 cluster boxes at once (a single group needs no box: the parent tested the leaf box)
        u32 g = gbase + lane;
        bool pass = false;
        if (g < ngroups)
        {
            if (ngroups == 1) pass = true;
            else
            {
                const float4* cp = tv.clusters + CLUSTER_F4 * (size_t)(lcf + g);
                float4 b0 = __ldg(cp), b1 = __ldg(cp + 1);
                boxes++;
                pass = cluster_hit<MODE>(mk3(b0.x, b0.y, b0.z), mk3(b0.w, b1.x, b1.y), ric, roi, rbest);
            }
        }
        unsigned bits = __ballot_sync(FULL, pass);
        // faces of the passing groups, 8 groups (32 faces) per step
        u32 npass = __popc(bits);
        for (u32 k0 = 0; k0 < npass; k0 += 32 / CLUSTER)
        {
            u32 k = k0 + lane / CLUSTER;
            if (k < npass)
            {
                u32 grp = gbase + __fns(bits, 0, (int)k + 1);
                u32 i = grp * CLUSTER + lane % CLUSTER;
                if (i < lcnt)
                {
                    const float4* p = tv.ltris + 3 * (size_t)(lfirst + i);
                    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
                    tests++;
                    test_face<MODE>(a, b, c, ro, rd, mine);
                }
            }
        }
    }
    // winner = min over lanes of (d, id)
    u32 khi = orderable(mine.t), klo = mine.tri;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
    {
        u32 ohi = __shfl_xor_sync(FULL, khi, m), olo = __shfl_xor_sync(FULL, klo, m);
        if (ohi < khi || (ohi == khi && olo < klo)) { khi = ohi; klo = olo; }
    }
    unsigned who = __ballot_sync(FULL, orderable(mine.t) == khi && mine.tri == klo);
    int w = __ffs(who) - 1;
    float wt = __shfl_sync(FULL, mine.t, w), wu = __shfl_sync(FULL, mine.u, w), wv = __shfl_sync(FULL, mine.v, w);
    u32 wtri = __shfl_sync(FULL, mine.tri, w), wmesh = __shfl_sync(FULL, mine.mesh, w), wface = __shfl_sync(FULL, mine.face, w);
    if (COUNT)
    {
        tests = __reduce_add_sync(FULL, tests); boxes = __reduce_add_sync(FULL, boxes);
    }
    if ((int)lane == owner)
    {
        if (COUNT) { nt += tests; nc += boxes; }
        if (wt < best.t || (wt == best.t && wt != FLT_MAX && wtri < best.tri))
        {
            best.t = wt; best.u = wu; best.v = wv; best.tri = wtri; best.mesh = wmesh; best.face = wface;
        }
    }
}

// Slot -> ray index.  With a width, slots enumerate 8x4 pixel tiles (one tile per warp-sized fetch) so that the lanes of a
// warp start on neighbouring pixels in both directions; slots that fall outside the image map to "no ray".
__device__ __forceinline__ bool slot_to_ray(u32 slot, u64 n, u32 width, u32 height, u64& ray)
{
    if (width == 0) { ray = slot; return slot < n; }
    u32 tilesX = (width + 7) >> 3;
    u32 tile = slot >> 5, in = slot & 31;
    u32 px = (tile % tilesX) * 8 + (in & 7), py = (tile / tilesX) * 4 + (in >> 3);
    ray = (u64)py * width + px;
    return px < width && py < height;
}

// cost -> one of TS_BUCKETS classes, two per power of two
__device__ __forceinline__ u32 tile_bucket(u32 cost)
{
    u32 l = 31u - (u32)__clz((int)(cost | 1u));
    u32 b = 2u * l + (l ? ((cost >> (l - 1u)) & 1u) : 0u);
    return b < (u32)TS_BUCKETS ? b : (u32)TS_BUCKETS - 1u;
}

// After a camera launch: histogram of the tiles' cost buckets ...
// filter (nullable): only tiles with filter[i] != 0 take part (bounce launches: the tiles in which the camera launch had hits)
__global__ void tile_hist_kernel(u32 tiles, const u32* __restrict__ cost, u32* state, const u32* __restrict__ filter)
{
    __shared__ u32 h[TS_BUCKETS];
    if (threadIdx.x < TS_BUCKETS) h[threadIdx.x] = 0;
    __syncthreads();
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < tiles && (!filter || filter[i] != 0u);
    u32 c = in ? cost[i] : 0u;
    if (in) atomicAdd(&h[tile_bucket(c)], 1u);
    // total cost of the frame (64 bits in state[TS_SUM], state[TS_SUM + 1])
    unsigned long long sum = c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(reinterpret_cast<unsigned long long*>(state + TS_SUM), sum);
    __syncthreads();
    if (threadIdx.x < TS_BUCKETS && h[threadIdx.x]) atomicAdd(&state[threadIdx.x], h[threadIdx.x]);
}

// ... then the work list of the next launch: tiles by descending cost bucket (counting sort; the order inside a bucket is arbitrary).
// Tiles that would be the launch's critical path -- cost above splitPct % of the ideal makespan (total cost / resident warps); on the
// bench frame some 250 silhouette tiles that each cost ~20x the average -- are entered as four QUARTERS (one 8-pixel row each, entry =
// tile | (row + 1) << 28), so that four warps share them.  Quarters run at 8 of 32 lanes, so splitting more than those loses throughput
// (measured: splitting the top 1/256 of the tiles -9 % launch time, the top 1/64 +0 %, the top 1/16 +27 %).  state[TS_ENTRIES] = number of entries; state[TS_OK] says the list is usable (every tile was costed).
// samples: a tile's cost is the sum over the launch's samples, but each sample of it is a work item of its own
__global__ void tile_order_kernel(u32 tiles, u32* __restrict__ cost, u32* state, u32* __restrict__ order, u32 splitPct, u32 residentWarps, const u32* __restrict__ filter, u32 samples)
{
    __shared__ u32 base[TS_BUCKETS], bins[TS_BUCKETS];
    __shared__ u32 total, entries, splitFrom;
    __shared__ unsigned long long costSum;
    if (threadIdx.x < TS_BUCKETS) bins[threadIdx.x] = state[threadIdx.x];      // one parallel read: the serial part below runs on shared memory
    if (threadIdx.x == TS_BUCKETS) costSum = *reinterpret_cast<const unsigned long long*>(state + TS_SUM);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        u32 run = 0, sf = TS_BUCKETS;
        if (splitPct)
        {
            unsigned long long thr = costSum / (residentWarps ? residentWarps : 1u) * splitPct / 100ull * (samples ? samples : 1u);    // cost a tile must exceed to be split
            for (int b = TS_BUCKETS - 1; b >= 2; --b)
            {
                u32 l = (u32)b >> 1;
                unsigned long long lower = (1ull << l) + (((u32)b & 1u) ? (1ull << (l - 1u)) : 0ull);     // smallest cost of bucket b
                if (lower >= thr && thr > 0 && run + bins[b] <= tiles / 32u) { run += bins[b]; sf = (u32)b; } else break;
            }
        }
        u32 pos = 0; run = 0;
        for (int b = TS_BUCKETS - 1; b >= 0; --b) { base[b] = pos; pos += bins[b] * ((u32)b >= sf ? 4u : 1u); run += bins[b]; }
        total = run; entries = pos; splitFrom = sf;
    }
    __syncthreads();
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    // this kernel is the last reader of the costs: it leaves them zeroed for the launch that accumulates the next ones
    const u32 myCost = i < tiles ? cost[i] : 0u;
    if (i < tiles) cost[i] = 0u;
    // without a filter the list is usable only if every tile was costed by the previous launch; a filtered list holds exactly the tiles
    // that passed the filter (a tile that has no cost yet lands in the cheapest class)
    if (i < tiles && (filter ? filter[i] != 0u : total == tiles))
    {
        u32 b = tile_bucket(myCost);
        // one atomic per (warp, bucket): neighbouring tiles mostly share a bucket
        unsigned live = __activemask(), peers = __match_any_sync(live, b), lane = threadIdx.x & 31u;
        unsigned leader = __ffs(peers) - 1u;
        u32 k = 0;
        if (lane == leader) k = atomicAdd(&state[TS_BUCKETS + b], (u32)__popc(peers));
        k = __shfl_sync(peers, k, (int)leader) + (u32)__popc(peers & ((1u << lane) - 1u));
        if (b >= splitFrom)
        {
            u32 pos = base[b] + 4u * k;
            if (pos + 3u < 4u * tiles) { order[pos] = i | (1u << 28); order[pos + 1] = i | (2u << 28); order[pos + 2] = i | (3u << 28); order[pos + 3] = i | (4u << 28); }
        }
        else
        {
            u32 pos = base[b] + k;
            if (pos < 4u * tiles) order[pos] = i;
        }
    }
    if (i == 0) { state[TS_ENTRIES] = entries; state[TS_OK] = (filter || total == tiles) ? 1u : 0u; }
}

// Persistent warps (one resident grid; every warp pulls work from a global counter), while-while traversal:
//   refill idle lanes -> all lanes with a pending internal node walk nodes until none has one -> all lanes with a pending
//   leaf test its triangles together -> lanes whose ray is finished write their result and become idle.
// Keeping the two phases warp-synchronous is what lifts SIMT efficiency: in the first version (profiles/r1_a) only 11 of
// 32 lanes were active on average because leaf loops and node walks of different lanes serialised each other.
#ifndef BM_TR_BLOCKS
#define BM_TR_BLOCKS 6
#endif
// SEG (ray batches through bricks, bricks.cu): a slot is one SEGMENT -- (ray, brick, position of the brick in the ray's own visiting order) --
// of a list sorted by brick; the lane walks the brick's subtree only and merges what it finds into the ray's 64-bit result key with
// an atomic minimum (SegIn).  Everything between refill and retire is the same traversal.
// BOUNCE (bounce launches, BounceIn): a camera launch whose lanes trace the diffuse bounce ray of their pixel's primary hit.
// samples > 1 (camera launches): the slots enumerate `samples` consecutive samples of every pixel, tile by tile.
template <int MODE, bool CAMERA, bool COUNT, bool WIDE, bool SEG, bool BOUNCE>
__global__ void __launch_bounds__(TR_THREADS, BM_TR_BLOCKS) trace_kernel(TreeView tv, const bmRay* __restrict__ rays, CameraIn cam, u64 n, u32 width, u32 height,
                                                           u32 totalSlots, const u32* __restrict__ numRaysDev, bmHit* __restrict__ hits, u32* __restrict__ pixels,
                                                           const bmStaticMeshData* __restrict__ meshData, u64* counters, u32* workCounter, int leafBreak, int leafWait, int refillMin, int queueDepth, TileSched sched,
                                                           SegIn seg, u32 samples, BounceIn bnc)
{
    const u32 rootRef = tv.numLeaves >= 2 ? __ldg(tv.root) : (tv.numLeaves == 1 ? (0u | LEAF_BIT) : REF_EMPTY);
    if (numRaysDev) { n = __ldg(numRaysDev); totalSlots = (u32)n; }       // count produced on the device (bmBounceRays): no host round trip
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31;
    __shared__ u32 pendS[CAMERA ? 1 : TR_QUEUE_MAX][TR_THREADS];      // postponed leaves of ray batches, [entry][thread]
#if BM_TMA_LEAF
    __shared__ __align__(128) unsigned char stageS[TR_THREADS / 32][TMA_WARP_BYTES];
    __shared__ __align__(8) unsigned long long barS[TR_THREADS / 32];
    u32 barPhase = 0;
    if ((threadIdx.x & 31) == 0) { mbar_init(&barS[threadIdx.x >> 5], 1u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
#endif
    u32 stack[TR_STACK];
    int top = -1;
    u32 cur = REF_EMPTY, npend = 0;
    bool active = false, exhausted = false;
    u64 rayIdx = 0;
    V3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), inv = mk3(0, 0, 1), ic = mk3(0, 0, 1), oi = mk3(0, 0, 0), fm = mk3(0, 0, 0);
    HitRec best = { FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };
    u32 nv = 0, nt = 0, nc = 0, rayStart = 0, longest = 0;      // COUNT only: totals, and the longest single ray (the kernel's critical path)
    // temporal tile scheduling (camera launches; see TileSched): the tile this warp is working on and when it started
    const bool useOrder = CAMERA && sched.order != nullptr && sched.state[TS_OK] != 0u;
    const u32 sampleSlots = (CAMERA || BOUNCE) ? totalSlots / samples : totalSlots;      // slots of one sample (camera launches without a schedule)
    if (useOrder) totalSlots = sched.state[TS_ENTRIES] * 32u * samples;      // the work list may hold some tiles as four quarter entries
    u32 raySample = 0;                      // CAMERA: which of the launch's samples this lane's ray belongs to
    u32 nBounce = 0;                        // BOUNCE: rays this lane has traced
    u32 curTile = 0xFFFFFFFFu;
    long long tileStart = 0;
    u32 segSeq = 0;                         // SEG: position of this brick in the ray's visiting order
    unsigned long long segKey0 = ~0ull;     // SEG: the ray's result key when the segment started

    for (;;)
    {
        // ---- refill -----------------------------------------------------------------------------------------------------
        unsigned idle = __ballot_sync(FULL, !active);
        if (CAMERA && sched.cost && idle == FULL && curTile != 0xFFFFFFFFu)
        {
            // the warp's tile is finished: its cost (32-cycle units) and cost bucket
            if (lane == 0)
            {
                unsigned long long c = (unsigned long long)(clock64() - tileStart) >> 5;
                u32 cost = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)c;
                atomicAdd(&sched.cost[curTile], cost);          // zeroed before the launch; the quarters of a split tile add up
            }

            curTile = 0xFFFFFFFFu;
        }
        if (idle && !exhausted && (__popc(idle) >= refillMin || idle == FULL))
        {
            u32 base = 0;
            if (lane == 0) base = atomicAdd(workCounter, (u32)__popc(idle));
            base = __shfl_sync(FULL, base, 0);
            if (base >= totalSlots) exhausted = true;
            u32 quarter = 0;                                     // 1..4: this fetch is one 8-pixel row of a split tile
            u32 fetchSample = 0;
            if (CAMERA && sched.cost && base < totalSlots)
            {
                // (scheduled camera launches refill whole warps only: idle == FULL and base is a multiple of 32)
                u32 entry = base >> 5;
                if (samples > 1u) { fetchSample = entry % samples; entry /= samples; }      // the samples of a tile follow each other: expensive tiles first for all of them
                if (useOrder) entry = __ldg(sched.order + entry);
                curTile = entry & 0x0FFFFFFFu; quarter = entry >> 28;
                tileStart = clock64();
            }
            if (!active)
            {
                u32 slot = base + __popc(idle & ((1u << lane) - 1u));
                bool take = slot < totalSlots;
                if (CAMERA && sched.cost && take)
                {
                    if (quarter) { take = lane < 8u; slot = (curTile << 5) | ((quarter - 1u) << 3) | (lane & 7u); }
                    else slot = (curTile << 5) | lane;
                }
                else if ((CAMERA || BOUNCE) && samples > 1u && take) { fetchSample = slot / sampleSlots; slot -= fetchSample * sampleSlots; }
                u32 startRef = rootRef;
                if (SEG && take)
                {
                    // the segment: its ray, its brick's subtree, and what the ray's other segments have found so far.  REFERENCE: the first
                    // brick (in the ray's own order) with a hit decides, so a segment behind a brick that already reported one is skipped;
                    // CLOSEST: the best hit so far bounds this segment's search.
                    const u32 key = __ldg(seg.keys + slot);
                    rayIdx = __ldg(seg.rays + slot);
                    segSeq = key >> 24;
                    startRef = __ldg(seg.roots + ((key >> seg.brickShift) & seg.brickMask));
                    segKey0 = __ldcg(seg.rayKey + rayIdx);
                    if (MODE == BM_TRACE_REFERENCE && (u32)(segKey0 >> 32) < segSeq) take = false;
                }
                if (take && !SEG) take = slot_to_ray(slot, n, width, height, rayIdx);
                float bt = 0.f; u32 btri = 0xFFFFFFFFu;
                if (BOUNCE && take)
                {
                    // the pixel's primary hit of this sample: no hit, no bounce ray
                    const float4 hf = __ldg(reinterpret_cast<const float4*>(bnc.hits1) + (size_t)fetchSample * n + rayIdx);
                    bt = hf.x; btri = __float_as_uint(hf.w);
                    take = btri != 0xFFFFFFFFu;
                }
                if (take)
                {
                    if (CAMERA || BOUNCE) raySample = fetchSample;
                    const u32 camSample = cam.sample == 0xFFFFFFFFu ? cam.sample : cam.sample + fetchSample;
                    if (BOUNCE)
                    {
                        V3 co, cd;
                        camera_ray(cam, rayIdx, camSample, co, cd);      // the primary ray exactly as its own launch formed it
                        bounce_ray(co, cd, bt, btri, bnc.soup, rayIdx, bnc.sample + fetchSample, bnc.bounce, bnc.seed, o, d);
                        nBounce++;
                    }
                    else if (CAMERA && !rays) camera_ray(cam, rayIdx, camSample, o, d);      // formed in registers: no ray record is ever written or read
                    else
                    {
                        const float* r = reinterpret_cast<const float*>(rays + rayIdx);
                        o = mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2));
                        d = mk3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
                    }
                    inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
                    ic = mk3(cull_inv(inv.x), cull_inv(inv.y), cull_inv(inv.z));
                    oi = o * ic;
                    if (WIDE && MODE == BM_TRACE_REFERENCE) fm = mk3(tv.leafExt[0] * fabsf(ic.x), tv.leafExt[1] * fabsf(ic.y), tv.leafExt[2] * fabsf(ic.z));
                    best.t = FLT_MAX; best.u = 0.f; best.v = 0.f; best.tri = best.mesh = best.face = 0xFFFFFFFFu;
                    if (SEG && MODE != BM_TRACE_REFERENCE && segKey0 != ~0ull) { best.t = __uint_as_float((u32)(segKey0 >> 32)); best.tri = (u32)segKey0; }
                    top = -1;
                    cur = startRef;
                    active = true;
                    if (COUNT) rayStart = nv + nt + nc;
                }
            }
        }
        // nothing to do: the work list is drained -- or the slots just fetched held no ray (a bounce launch's tile without primary hits,
        // segments another brick has already decided): fetch again
        if (!__any_sync(FULL, active)) { if (exhausted) break; continue; }

        // ---- phase 1: internal nodes ----------------------------------------------------------------------------------------------
        // Ray batches carry a small queue of POSTPONED leaves per lane (queueDepth entries in shared memory): a lane that reaches a leaf
        // notes it and walks on as if the leaf held no hit, and the noted leaves are tested together in phase 2, in the order they were
        // met.  Incoherent rays reach their leaves at very different times: with lanes waiting AT their leaf, the node walk of a
        // 20 M-triangle soup ran at 13 of 32 lanes (profiles/r2a).  Walking on is exact: what the traversal does after a leaf depends on
        // the leaf only through "stop here" (reference mode: the first leaf with a hit ends the ray, so at worst the nodes visited past it
        // were visited for nothing -- one leaf in ~15 of such a batch has a hit) or through the shrinking [0, best t] (closest mode: the
        // cull lags by the queue).  queueDepth = 0 (camera launches: coherent tiles arrive at their leaves together) is the lane-waits rule.
        int waited = 0;                                    // steps taken while at least leafBreak lanes could not move
        for (;;)
        {
            const bool hasCur = active && cur != REF_EMPTY;
            const bool inner = hasCur && !(cur & LEAF_BIT);
            const bool canq = !CAMERA && hasCur && (cur & LEAF_BIT) && npend < (u32)queueDepth;
            if (!__any_sync(FULL, inner || canq)) break;
            if (!CAMERA)
            {
                // lanes that cannot move before the leaf phase: at a leaf with a full queue, or out of nodes with leaves still noted;
                // lanes that are finished (or empty) and could take a new ray
                const unsigned stalled = __ballot_sync(FULL, active && !inner && !canq && (hasCur || npend > 0u));
                const unsigned spare = __ballot_sync(FULL, !active || (!hasCur && npend == 0u));
                if (__popc(stalled) >= leafBreak && ++waited > leafWait) break;
                if (queueDepth && !exhausted && __popc(spare) >= refillMin) break;
            }
            if (canq)
            {
                pendS[npend][threadIdx.x] = cur; npend++;
                cur = top >= 0 ? stack[top--] : REF_EMPTY;
            }
            else if (inner)
            {
                if (COUNT) nv++;
                if (WIDE) wide_node<MODE>(tv, cur, o, d, inv, ic, oi, fm, best.t, stack, top);
                else if (MODE == BM_TRACE_REFERENCE) ref_node(tv, cur, o, d, inv, ic, oi, stack, top);
                else closest_node(tv, cur, o, inv, best.t, stack, top);
                cur = top >= 0 ? stack[top--] : REF_EMPTY;
            }
        }

        // ---- phase 2: the noted leaves of every lane, then the leaf it stands at --------------------------------------------------------
        // stage A: every lane scans the cluster boxes of its leaf (groups of CLUSTER faces) and keeps a bit mask of the groups the
        // ray's line meets; stage B: lanes pop one group at a time and test its faces in lock step.  Both loops are warp-uniform.
        bool finished = false;
        const bool curLeaf = active && cur != REF_EMPTY && (cur & LEAF_BIT);
        const u32 nleaf = active ? npend + (curLeaf ? 1u : 0u) : 0u;
        const u32 rounds = __reduce_max_sync(FULL, nleaf);
        for (u32 round = 0; round < rounds; round++)
        {
            const bool isLeaf = !finished && round < nleaf;
            const u32 lref = (!CAMERA && round < npend) ? pendS[round][threadIdx.x] : cur;
            u32 first = 0, cnt = 0, cfirst = 0;
            if (isLeaf)
            {
                if (COUNT) nv++;
                const float4* lp = tv.leaves + LEAF_F4 * (size_t)(lref & ~LEAF_BIT);
                float4 r0 = __ldg(lp), r1 = __ldg(lp + 1), r2 = __ldg(lp + 2);
                first = __float_as_uint(r0.y); cfirst = __float_as_uint(r0.w);
                if (MODE == BM_TRACE_REFERENCE)
                {
                    if (box_ray(mk3(r1.x, r1.y, r1.z), mk3(r1.w, r2.x, r2.y), o, inv) != FLT_MAX) cnt = __float_as_uint(r0.z);      // BuildTree.cu:405-406
                }
                else
                {
                    // CLOSEST: a hit inside this leaf's cell (grown by cellPad) is the only kind this leaf has to find (see closest_node)
                    const float pad = tv.cellPad;
                    float tn;
                    if (box_interval(mk3(r1.x - pad, r1.y - pad, r1.z - pad), mk3(r1.w + pad, r2.x + pad, r2.y + pad), o, inv, best.t, tn)) cnt = __float_as_uint(r0.z);
                }
            }
            // few lanes with a sizeable leaf to test: let the whole warp work on them one after the other (small leaves are cheaper
            // on their own lane)
            unsigned pendAll = __ballot_sync(FULL, cnt > 0u);
            unsigned pend = __ballot_sync(FULL, cnt > (u32)COOP_MIN_FACES);
            if (pend && __popc(pendAll) <= COOP_MAX_LANES)
            {
                while (pend)
                {
                    int owner = __ffs(pend) - 1;
                    pend &= pend - 1;
#if BM_TMA_LEAF
                    coop_leaf_staged<MODE, COUNT>(tv, lane, owner, o, d, ic, oi, first, cnt, cfirst, best, nt, nc, stageS[threadIdx.x >> 5], &barS[threadIdx.x >> 5], barPhase);
#else
                    coop_leaf<MODE, COUNT>(tv, lane, owner, o, d, ic, oi, first, cnt, cfirst, best, nt, nc);
#endif
                }
                if (cnt > (u32)COOP_MIN_FACES) cnt = 0;   // done: skip the per-lane path below
            }
            u32 ncl = (cnt + CLUSTER - 1) / CLUSTER;
            u32 maxNcl = __reduce_max_sync(FULL, ncl);
            // 64 groups per round (one bit each); leaves above 256 faces simply take several rounds
            for (u32 cb = 0; cb < maxNcl; cb += 64)
            {
                u32 here = ncl > cb ? min(ncl - cb, 64u) : 0u;                        // groups of this lane in this round
                unsigned long long mask = (here == 1u && ncl == 1u) ? 1ull : 0ull;    // a single group: its box is the leaf box the parent tested
                u32 scan = (ncl > 1u) ? here : 0u;                                    // groups whose boxes this lane has to look at
                u32 maxScan = __reduce_max_sync(FULL, scan);
                {
                    // four boxes per trip with all eight loads issued before the first test: the tests are independent, so the L2
                    // latency is paid once per trip instead of once per box (the longest rays graze many populated leaves)
                    const float4* cp = tv.clusters + CLUSTER_F4 * (size_t)(cfirst + cb);
                    for (u32 i = 0; i < maxScan; i += 4, cp += 4 * CLUSTER_F4)
                    {
                        if (i < scan)
                        {
                            float4 b[8];
#pragma unroll
                            for (int k = 0; k < 4; k++)
                            {
                                bool ok = i + k < scan;
                                b[2 * k] = ok ? __ldg(cp + 2 * k) : make_float4(0, 0, 0, 0);
                                b[2 * k + 1] = ok ? __ldg(cp + 2 * k + 1) : make_float4(0, 0, 0, 0);
                            }
#pragma unroll
                            for (int k = 0; k < 4; k++)
                            {
                                if (i + k < scan)
                                {
                                    if (COUNT) nc++;
                                    if (cluster_hit<MODE>(mk3(b[2 * k].x, b[2 * k].y, b[2 * k].z), mk3(b[2 * k].w, b[2 * k + 1].x, b[2 * k + 1].y), ic, oi, best.t)) mask |= 1ull << (i + k);
                                }
                            }
                        }
                    }
                }
                while (__any_sync(FULL, mask != 0ull))
                {
                    if (mask)
                    {
                        u32 g = cb + (u32)__ffsll((long long)mask) - 1u;
                        mask &= mask - 1ull;
                        u32 i0 = g * CLUSTER, i1 = min(i0 + CLUSTER, cnt);
                        const float4* p = tv.ltris + 3 * (size_t)(first + i0);
                        // two faces per trip, their six loads issued together (same reasoning as the box scan)
                        for (u32 i = i0; i < i1; i += 2, p += 6)
                        {
                            bool two = i + 1 < i1;
                            float4 a0 = __ldg(p), b0 = __ldg(p + 1), c0 = __ldg(p + 2);
                            float4 a1 = two ? __ldg(p + 3) : a0, b1 = two ? __ldg(p + 4) : b0, c1 = two ? __ldg(p + 5) : c0;
                            if (COUNT) nt += two ? 2 : 1;
                            test_face<MODE>(a0, b0, c0, o, d, best);
                            if (two) test_face<MODE>(a1, b1, c1, o, d, best);
                        }
                    }
                }
            }
            if (isLeaf && MODE == BM_TRACE_REFERENCE && best.t != FLT_MAX) finished = true;          // :427-431: the first leaf with a hit ends the ray
        }
        if (curLeaf && !finished) cur = top >= 0 ? stack[top--] : REF_EMPTY;
        npend = 0;

        // ---- retire ------------------------------------------------------------------------------------------------------------
        if (active && (finished || cur == REF_EMPTY))
        {
            if (SEG)
            {
                // REFERENCE: (position of the brick in the ray's order, triangle) -- the smallest position wins; CLOSEST: (t >= 0 as its
                // bit pattern, triangle) -- the smallest t wins, the smaller id on a tie.  t, u, v are recomputed from the triangle afterwards.
                if (best.tri != 0xFFFFFFFFu)
                {
                    const unsigned long long k = MODE == BM_TRACE_REFERENCE ? (((unsigned long long)segSeq << 32) | best.tri)
                                                                            : (((unsigned long long)__float_as_uint(best.t == 0.f ? 0.f : best.t) << 32) | best.tri);
                    if (k < segKey0) atomicMin(seg.rayKey + rayIdx, k);
                }
            }
            else
            {
                const size_t at = (CAMERA || BOUNCE) ? (size_t)raySample * n + rayIdx : (size_t)rayIdx;
                // tiles with at least one hit (in any sample), for the bounce launch that follows: its work list leaves the others out
                // (the lanes of a warp share the tile: one store per warp instruction)
                if (CAMERA && !BOUNCE && sched.hitFlag && best.tri != 0xFFFFFFFFu) sched.hitFlag[curTile] = 1u;
                if (hits) { reinterpret_cast<float4*>(hits)[at] = make_float4(best.t, best.u, best.v, __uint_as_float(best.tri)); }
                if (pixels) pixels[rayIdx] = shade(best, meshData);
                if (BOUNCE)
                {
                    // the progressive estimator, applied where the ray ends: this sample of the pixel had a primary hit (ambient) whose bounce
                    // ray escaped (+ weight) or not.  Several samples of a pixel may retire at once: atomic (the terms are exact in binary)
                    if (bnc.accum) atomicAdd(bnc.accum + rayIdx, bnc.ambient + (best.tri == 0xFFFFFFFFu ? bnc.weight : 0.f));
                    if (bnc.covered) bnc.covered[rayIdx] = 1u;
                }
            }
            active = false;
            if (COUNT) longest = max(longest, nv + nt + nc - rayStart);
        }
    }
    if (BOUNCE && bnc.rayCount)
    {
        nBounce = __reduce_add_sync(FULL, nBounce);
        if (lane == 0 && nBounce) atomicAdd(bnc.rayCount, (unsigned long long)nBounce);
    }
    if (COUNT)
    {
        u64 a = nv, b = nt, c = nc;
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) { a += __shfl_down_sync(FULL, a, k); b += __shfl_down_sync(FULL, b, k); c += __shfl_down_sync(FULL, c, k); }
        if (lane == 0)
        {
            atomicAdd((unsigned long long*)&counters[0], (unsigned long long)a); atomicAdd((unsigned long long*)&counters[1], (unsigned long long)b);
            atomicAdd((unsigned long long*)&counters[2], (unsigned long long)c);
        }
        longest = __reduce_max_sync(FULL, longest);
        if (lane == 0) atomicMax((unsigned long long*)&counters[3], (unsigned long long)longest);
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

// leaves a lane of a ray batch postpones before it has to wait for the leaf phase (BM_LEAF_QUEUE overrides, 0 = wait at the first leaf)
int queue_depth()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("BM_LEAF_QUEUE"); v = e ? atoi(e) : 4; if (v < 0) v = 0; if (v > TR_QUEUE_MAX) v = TR_QUEUE_MAX; }
    return v;
}

// phase 1 hands over to phase 2 once this many lanes of a warp cannot move (BM_LEAF_BREAK overrides, for tuning): with lanes waiting at
// their first leaf (queue depth 0) half the warp, with a leaf queue a quarter
int leaf_break(int queueDepth)
{
    static int v = 0;
    if (!v) { const char* e = getenv("BM_LEAF_BREAK"); v = e ? atoi(e) : -1; if (v == 0) v = 1; if (v > 33) v = 33; }
    return v > 0 ? v : (queueDepth ? 8 : 16);
}

int leaf_wait(int queueDepth)
{
    static int v = -2;
    if (v == -2) { const char* e = getenv("BM_LEAF_WAIT"); v = e ? atoi(e) : -1; }
    return v >= 0 ? v : (queueDepth ? 2 : 6);
}

// A warp refills only when this many of its lanes are idle.  Measured: camera launches are 5 % faster at 32 (a warp finishes its whole
// 8x4 tile before taking the next one: keeping the lanes on neighbouring rays is worth more than filling the holes early); ray batches
// want 8 (an incoherent batch on the 10 M soup: 54.9 Mrays/s at 8 vs 38.9 at 32; the c2 bounce batch loses 7 % against 32).
// BM_REFILL_MIN overrides both, for tuning.
// a tile is split into quarters when it cost more than this percentage of the ideal makespan; BM_TILE_SPLIT=0 disables splitting
u32 tile_split_pct()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("BM_TILE_SPLIT"); v = e ? atoi(e) : 75; if (v < 0) v = 75; }
    return (u32)v;
}

int refill_min(bool camera)
{
    static int v = 0;
    if (!v) { const char* e = getenv("BM_REFILL_MIN"); v = e ? atoi(e) : -1; if (v > 32) v = 32; if (v == 0) v = 1; }
    return v > 0 ? v : (camera ? 32 : 8);
}

} // namespace
// camera launches run under a tile schedule only when they refill whole warps (the default; BM_REFILL_MIN overrides it for tuning)
bool tile_sched_usable() { return refill_min(true) == 32; }
namespace
{
// node records walked by the trace kernels (bmTreeNodeLayout)
bool use_wide(const TreeView& tv)
{
    if (!tv.nodes || !tv.wnodes) return tv.wnodes != nullptr;      // the tree holds one kind of record
    static int v = -1;                                              // BM_NODE_LAYOUT=both: BM_WIDE=0 walks the binary records (A/B runs)
    if (v < 0) { const char* e = getenv("BM_WIDE"); v = e ? atoi(e) : 1; }
    return v != 0;
}

template <int MODE, bool CAMERA, bool COUNT, bool WIDE, bool SEG = false, bool BOUNCE = false>
int launch_trace(const TreeView& tv, const bmRay* d_rays, const CameraIn& cam, u64 n, const u32* d_numRays, u32 width, bmHit* d_hits, u32* d_pixels,
                 const bmStaticMeshData* d_meshData, u64* d_counters, u32* d_workCounter, const TileSched* sched, cudaStream_t s, SegIn seg = SegIn{ nullptr, nullptr, nullptr, nullptr, 0u, 0u },
                 u32 samples = 1, const BounceIn* bounce = nullptr)
{
    static std::atomic<int> occ[MAX_DEVICES];          // per device and per instantiation
    const int dev = current_device();
    int blocksPerSM = occ[dev].load(std::memory_order_relaxed);
    if (!blocksPerSM)
    {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, trace_kernel<MODE, CAMERA, COUNT, WIDE, SEG, BOUNCE>, TR_THREADS, 0) != cudaSuccess || blocksPerSM < 1) blocksPerSM = 4;
        occ[dev].store(blocksPerSM, std::memory_order_relaxed);
    }
    u32 height = width ? (u32)((n + width - 1) / width) : 0;
    const u64 sampleSlots = width ? (u64)((width + 7) / 8) * ((height + 3) / 4) * 32 : n;
    if (!(CAMERA || BOUNCE) || !samples) samples = 1;
    u64 slots = sampleSlots * samples;
    if ((samples > 1 ? slots * 4 : slots) >= 0xFFFFFFF0ull) { set_error("more than 2^32 rays in one launch"); return BM_ERROR_CAPACITY; }       // a schedule may list a tile as four quarter entries
    u64 warps = (slots + 31) / 32;
    u64 blocks = (warps + TR_THREADS / 32 - 1) / (TR_THREADS / 32);
    u64 cap = (u64)sm_count() * blocksPerSM;
    if (blocks > cap) blocks = cap;
    TileSched ts = { nullptr, nullptr, nullptr, 0 };
    const int refill = refill_min(CAMERA);
    if (CAMERA && sched && !COUNT && refill == 32 && sched->tiles == (u32)(sampleSlots / 32))
    {
        ts = *sched;
        BM_CUDA(cudaMemsetAsync(ts.state, 0, (TS_WORK + 1) * 4, s));          // histogram, cursors, cost sum and the work counter; OK flag and entry count of the previous launch stay
        d_workCounter = ts.state + TS_WORK;
        if (ts.filter)
        {
            // a filtered work list is made just before the launch it serves (the filter is what the preceding camera launch found out):
            // tiles that passed, by descending cost in the previous launch of this kind
            tile_hist_kernel<<<div_up(ts.tiles, 256), 256, 0, s>>>(ts.tiles, ts.cost, ts.state, ts.filter);
            BM_LAUNCHED();
            tile_order_kernel<<<div_up(ts.tiles, 256), 256, 0, s>>>(ts.tiles, ts.cost, ts.state, ts.order, tile_split_pct(), (u32)blocks * (TR_THREADS / 32), ts.filter, samples);
            BM_LAUNCHED();
        }
        if (ts.hitFlag) BM_CUDA(cudaMemsetAsync(ts.hitFlag, 0, (size_t)ts.tiles * 4, s));
    }
    else BM_CUDA(cudaMemsetAsync(d_workCounter, 0, 4, s));
    const BounceIn none = { nullptr, nullptr, 0u, 0u, 0u, nullptr, nullptr, 0.f, 0.f, nullptr };
    trace_kernel<MODE, CAMERA, COUNT, WIDE, SEG, BOUNCE><<<(unsigned)blocks, TR_THREADS, 0, s>>>(tv, d_rays, cam, n, width, height, (u32)slots, d_numRays, d_hits, d_pixels, d_meshData,
                                                                            d_counters, d_workCounter, leaf_break(CAMERA ? 0 : queue_depth()), leaf_wait(CAMERA ? 0 : queue_depth()), refill, CAMERA ? 0 : queue_depth(), ts, seg,
                                                                            samples, bounce ? *bounce : none);
    BM_LAUNCHED();
    if (ts.cost && !ts.filter)
    {
        tile_hist_kernel<<<div_up(ts.tiles, 256), 256, 0, s>>>(ts.tiles, ts.cost, ts.state, nullptr);
        BM_LAUNCHED();
        tile_order_kernel<<<div_up(ts.tiles, 256), 256, 0, s>>>(ts.tiles, ts.cost, ts.state, ts.order, tile_split_pct(), (u32)blocks * (TR_THREADS / 32), nullptr, samples);
        BM_LAUNCHED();
    }
    return BM_OK;
}

template <bool CAMERA, bool WIDE>
int dispatch_trace2(const TreeView& tv, const bmRay* d_rays, const CameraIn& cam, u64 n, const u32* d_numRays, u32 width, int mode, bmHit* d_hits, u32* d_pixels,
                    const bmStaticMeshData* d_meshData, u64* d_counters, u32* d_workCounter, const TileSched* sched, cudaStream_t s)
{
    if (mode == BM_TRACE_REFERENCE)
        return d_counters ? launch_trace<BM_TRACE_REFERENCE, CAMERA, true, WIDE>(tv, d_rays, cam, n, d_numRays, width, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s)
                          : launch_trace<BM_TRACE_REFERENCE, CAMERA, false, WIDE>(tv, d_rays, cam, n, d_numRays, width, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s);
    return d_counters ? launch_trace<BM_TRACE_CLOSEST, CAMERA, true, WIDE>(tv, d_rays, cam, n, d_numRays, width, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s)
                      : launch_trace<BM_TRACE_CLOSEST, CAMERA, false, WIDE>(tv, d_rays, cam, n, d_numRays, width, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s);
}

template <bool CAMERA>
int dispatch_trace(const TreeView& tv, const bmRay* d_rays, const CameraIn& cam, u64 n, const u32* d_numRays, u32 width, int mode, bmHit* d_hits, u32* d_pixels,
                   const bmStaticMeshData* d_meshData, u64* d_counters, u32* d_workCounter, const TileSched* sched, cudaStream_t s)
{
    if (use_wide(tv)) return dispatch_trace2<CAMERA, true>(tv, d_rays, cam, n, d_numRays, width, mode, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s);
    return dispatch_trace2<CAMERA, false>(tv, d_rays, cam, n, d_numRays, width, mode, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s);
}
} // namespace

int trace_rays(const TreeView& tv, const bmRay* d_rays, u64 n, const u32* d_numRays, int mode, bmHit* d_hits, u64* d_counters, u32* d_workCounter, cudaStream_t s)
{
    if (!n) return BM_OK;
    CameraIn cam = make_camera(nullptr, nullptr, nullptr);
    return dispatch_trace<false>(tv, d_rays, cam, n, d_numRays, 0, mode, d_hits, nullptr, nullptr, d_counters, d_workCounter, nullptr, s);
}

// the segments of a ray batch (bricks.cu): numSegs slots, each a (ray, brick) pair; results are merged into seg.rayKey
int trace_segments(const TreeView& tv, const bmRay* d_rays, u64 numSegs, int mode, const SegIn& seg, u64* d_counters, u32* d_workCounter, cudaStream_t s)
{
    if (!numSegs) return BM_OK;
    CameraIn cam = make_camera(nullptr, nullptr, nullptr);
    const bool wide = use_wide(tv);
#define BM_SEG_LAUNCH(M, C, W) launch_trace<M, false, C, W, true>(tv, d_rays, cam, numSegs, nullptr, 0, nullptr, nullptr, nullptr, d_counters, d_workCounter, nullptr, s, seg)
    if (mode == BM_TRACE_REFERENCE)
    {
        if (wide) return d_counters ? BM_SEG_LAUNCH(BM_TRACE_REFERENCE, true, true) : BM_SEG_LAUNCH(BM_TRACE_REFERENCE, false, true);
        return d_counters ? BM_SEG_LAUNCH(BM_TRACE_REFERENCE, true, false) : BM_SEG_LAUNCH(BM_TRACE_REFERENCE, false, false);
    }
    if (wide) return d_counters ? BM_SEG_LAUNCH(BM_TRACE_CLOSEST, true, true) : BM_SEG_LAUNCH(BM_TRACE_CLOSEST, false, true);
    return d_counters ? BM_SEG_LAUNCH(BM_TRACE_CLOSEST, true, false) : BM_SEG_LAUNCH(BM_TRACE_CLOSEST, false, false);
#undef BM_SEG_LAUNCH
}

int trace_camera(const TreeView& tv, const CameraIn& cam, u64 n, u32 width, int mode, bmHit* d_hits, u32* d_pixels, const bmStaticMeshData* d_meshData,
                 u64* d_counters, u32* d_workCounter, const TileSched* sched, cudaStream_t s, u32 samples, const BounceIn* bounce)
{
    if (!n) return BM_OK;
    if (samples > 1u && (cam.sample == 0xFFFFFFFFu || d_pixels)) { set_error("a multi-sample camera launch needs jittered samples and writes no pixels"); return BM_ERROR_INVALID_PARAMETER; }
    const SegIn noSeg = SegIn{ nullptr, nullptr, nullptr, nullptr, 0u, 0u };
    const bool wide = use_wide(tv);
#define BM_CAM_LAUNCH(M, C, W, B) launch_trace<M, true, C, W, false, B>(tv, nullptr, cam, n, nullptr, width, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s, noSeg, samples, bounce)
#define BM_BATCH_BOUNCE(C, W) launch_trace<BM_TRACE_CLOSEST, false, C, W, false, true>(tv, nullptr, cam, n, nullptr, width, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, nullptr, s, noSeg, samples, bounce)
    if (bounce)
    {
        // bounce launches are closest-hit launches (t >= 0: in the reference's any-sign semantics a bounce ray "hits" its own source triangle)
        if (mode != BM_TRACE_CLOSEST) { set_error("bounce launches trace closest hits"); return BM_ERROR_INVALID_PARAMETER; }
        // Scheduling: the slots are the pixels in 8x4 tile order either way.  "tiles" (default): whole tiles per warp under the launch's own
        // temporal tile schedule, as camera launches; "batch" (BM_BOUNCE_STYLE=batch): lanes refill as they fall idle and postpone leaves, as
        // ray batches do.  Measured (B200, frame of c2 / c3 / c4): tiles 0.864 / 13.48 / 55.1 ms, batch 0.910 / 13.64 / 55.6 ms -- in batch
        // style every warp pays a work-counter round trip per 8 slots of background, which costs more than the idle lanes of a tile.
        static const bool tiles = !(getenv("BM_BOUNCE_STYLE") && !strcmp(getenv("BM_BOUNCE_STYLE"), "batch"));
        if (!tiles)
        {
            if (wide) return d_counters ? BM_BATCH_BOUNCE(true, true) : BM_BATCH_BOUNCE(false, true);
            return d_counters ? BM_BATCH_BOUNCE(true, false) : BM_BATCH_BOUNCE(false, false);
        }
        if (wide) return d_counters ? BM_CAM_LAUNCH(BM_TRACE_CLOSEST, true, true, true) : BM_CAM_LAUNCH(BM_TRACE_CLOSEST, false, true, true);
        return d_counters ? BM_CAM_LAUNCH(BM_TRACE_CLOSEST, true, false, true) : BM_CAM_LAUNCH(BM_TRACE_CLOSEST, false, false, true);
    }
    if (samples > 1u)
    {
        if (mode == BM_TRACE_REFERENCE)
        {
            if (wide) return d_counters ? BM_CAM_LAUNCH(BM_TRACE_REFERENCE, true, true, false) : BM_CAM_LAUNCH(BM_TRACE_REFERENCE, false, true, false);
            return d_counters ? BM_CAM_LAUNCH(BM_TRACE_REFERENCE, true, false, false) : BM_CAM_LAUNCH(BM_TRACE_REFERENCE, false, false, false);
        }
        if (wide) return d_counters ? BM_CAM_LAUNCH(BM_TRACE_CLOSEST, true, true, false) : BM_CAM_LAUNCH(BM_TRACE_CLOSEST, false, true, false);
        return d_counters ? BM_CAM_LAUNCH(BM_TRACE_CLOSEST, true, false, false) : BM_CAM_LAUNCH(BM_TRACE_CLOSEST, false, false, false);
    }
#undef BM_CAM_LAUNCH
#undef BM_BATCH_BOUNCE
    return dispatch_trace<true>(tv, nullptr, cam, n, nullptr, width, mode, d_hits, d_pixels, d_meshData, d_counters, d_workCounter, sched, s);
}

int shade_hits(const bmHit* d_hits, u64 n, const float4* d_soup, const bmStaticMeshData* d_meshData, const bmMaterial* d_materials, const bmTexture* d_textures,
               u32 numTextures, u32* d_pixels, cudaStream_t s)
{
    if (!n) return BM_OK;
    shade_hits_kernel<<<div_up(n, 256), 256, 0, s>>>(d_hits, n, d_soup, d_meshData, d_materials, d_textures, numTextures, d_pixels);
    BM_LAUNCHED();
    return BM_OK;
}

// a ray batch that IS an image (width x n/width rays in row-major order, e.g. jittered camera rays): traced with the camera launches'
// scheduling -- 8x4 tiles, whole-warp refills, temporal tile schedule -- instead of the generic batch scheduling
int trace_rays_tiled(const TreeView& tv, const bmRay* d_rays, u64 n, u32 width, int mode, bmHit* d_hits, u64* d_counters, u32* d_workCounter,
                     const TileSched* sched, cudaStream_t s)
{
    if (!n) return BM_OK;
    CameraIn cam = make_camera(nullptr, nullptr, nullptr);
    return dispatch_trace<true>(tv, d_rays, cam, n, nullptr, width, mode, d_hits, nullptr, nullptr, d_counters, d_workCounter, sched, s);
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
