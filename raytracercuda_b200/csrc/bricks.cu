// bricks.cu -- ray batches against trees far larger than the L2, traced brick by brick.
//
// The reference's kd-tree splits the world box at the midpoint on x, y, z, x, ... (BuildTree.cu:196-216), so its top B levels are an
// implicit grid of 2^B "bricks" whatever the scene is, and what its traversal does up there (BuildTree.cu:394-481: cell slab test,
// children ordered by the side of the split plane on which the ray enters the cell) is a pure function of the ray.  An incoherent batch
// walks a tree of tens of GB with one DRAM round trip per node; cut into bricks, the same work runs out of the cache:
//   1. brick_table_kernel    : brick -> reference of the subtree that holds exactly the leaves whose code starts with the brick's bits
//                              (the radix node covering that code range: found by binary search, no build-side bookkeeping, so it also
//                              works on an imported replica), and the occupancy of the implicit levels above;
//   2. brick_segments_kernel : every ray walks the implicit top levels exactly as the reference would and notes the bricks it enters, in
//                              order: one SEGMENT (position in that order, brick, ray) each;
//   3. radix sort by brick (sort.cu), then trace_kernel<SEG> (trace.cu) walks every segment's subtree -- consecutive warps work on the
//      same few bricks -- and merges hits into one 64-bit key per ray with an atomic minimum:
//         REFERENCE: key = position << 32 | triangle   (the reference stops at the first leaf with a hit; leaves of an earlier brick come
//                                                       first, and inside a brick the segment's own walk stops at its first such leaf)
//         CLOSEST  : key = bits(t) << 32 | triangle    (t >= 0; the smaller id on a tie, as in the direct traversal)
//   4. brick_resolve_kernel  : key -> hit record; t, u, v recomputed from the triangle (same operations on the same operands as the leaf
//      loop, hence the same bits).
#include "bm_device.cuh"

namespace bm
{
namespace
{
constexpr int BR_MAX_SEGS = 3 * 32 + 4;       // a line crosses at most 3 * 2^(B/3) - 2 cells of the brick grid (B <= 15)

__device__ __forceinline__ u32 leaf_code(const TreeView& tv, u32 i) { return __float_as_uint(__ldg(tv.leaves + LEAF_F4 * (size_t)i).x); }

// first leaf whose code is >= key (codes are sorted, D-bit); key may be 2^D
__device__ u32 lower_bound_code(const TreeView& tv, u32 lo, u32 hi, u64 key)
{
    while (lo < hi)
    {
        u32 mid = lo + ((hi - lo) >> 1);
        if ((u64)leaf_code(tv, mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void brick_table_kernel(TreeView tv, int bits, u32* __restrict__ roots, unsigned char* __restrict__ occ)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (1u << bits)) return;
    const int D = tv.grid.D, sh = D - bits;
    const u32 lo = lower_bound_code(tv, 0, tv.numLeaves, (u64)b << sh);
    const u32 end = lower_bound_code(tv, lo, tv.numLeaves, ((u64)b + 1ull) << sh);
    u32 ref = BRICK_EMPTY;
    if (end == lo + 1) ref = lo | LEAF_BIT;
    else if (end > lo + 1)
    {
        // the radix node covering leaves [lo, end): its index is its split position (hierarchy.cu: parent = last leaf of the left child),
        // i.e. the last leaf whose code has a 0 at the highest bit in which the range's first and last code differ
        const u32 cl = leaf_code(tv, lo), cr = leaf_code(tv, end - 1);
        const int hb = 31 - __clz((int)(cl ^ cr));
        ref = lower_bound_code(tv, lo, end, (u64)((cr >> hb) << hb)) - 1u;
    }
    roots[b] = ref;
    occ[(1u << bits) + b] = ref != BRICK_EMPTY;
}

// occupancy of the implicit levels above the bricks, one level per launch (a level has at most 2^14 nodes)
__global__ void brick_occ_kernel(int depth, unsigned char* occ)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << depth)) return;
    const u32 h = (1u << depth) + i;
    occ[h] = occ[2 * h] | occ[2 * h + 1];
}

// The implicit top of the reference's traversal for one ray (oracle.cpp march_ray / BuildTree.cu:394-481 restricted to depth <= B):
// pop a node, test its cell with the reference's slab arithmetic, order the two children by the side of the split plane on which the ray
// enters the cell, push far then near; a node at depth B is a brick and is noted instead.  Children without any leaf are not pushed
// (the reference has no such node).  CLOSEST: the direct traversal culls with the boxes of whole triangles, which reach beyond their
// cells by at most a triangle; a hit point itself always lies in a cell of its triangle, up to rounding, so here the cells are grown by
// `pad` and every brick the forward ray touches is noted (the order only affects how early the search interval shrinks).
template <int MODE>
__global__ void __launch_bounds__(256) brick_segments_kernel(Grid g, int bits, const u32* __restrict__ roots, const unsigned char* __restrict__ occ,
                                                              const bmRay* __restrict__ rays, u64 n, float pad, u32* __restrict__ keys, u32* __restrict__ rayIdx,
                                                              u32 capacity, unsigned long long* count)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned short list[BR_MAX_SEGS];
    u32 cnt = 0;
    if (i < n)
    {
        const float* r = reinterpret_cast<const float*>(rays + i);
        const V3 o = mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2)), d = mk3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
        const V3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
        u32 stack[BRICK_MAX_BITS + 2];
        int top = 0;
        stack[0] = 0u;                                    // depth << 16 | path
        while (top >= 0)
        {
            const u32 e = stack[top--];
            const int depth = (int)(e >> 16);
            const u32 path = e & 0xFFFFu;
            u32 ix, iy, iz; decode_path(depth, path, ix, iy, iz);
            V3 mn, mx; cell_box(g, depth, ix, iy, iz, mn, mx);
            if (MODE != BM_TRACE_REFERENCE) { mn = mn - mk3(pad, pad, pad); mx = mx + mk3(pad, pad, pad); }
            const float boxDist = box_ray(mn, mx, o, inv);
            if (boxDist == FLT_MAX) continue;
            if (depth == bits) { if (cnt < (u32)BR_MAX_SEGS) list[cnt++] = (unsigned short)path; continue; }
            const int axis = depth % 3;
            const u32 ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
            const int sh = g.k[axis] - axis_bits(depth, axis);
            const float split = grid_at(g, axis, ((ia << 1) | 1u) << (sh - 1));
            const float p = get(o, axis) + boxDist * get(d, axis);
            const bool leftFirst = p < split;
            const u32 h = (2u << depth) + (path << 1);    // heap index of the lower child
            const u32 nearBit = leftFirst ? 0u : 1u, farBit = nearBit ^ 1u;
            if (occ[h + farBit]) stack[++top] = ((u32)(depth + 1) << 16) | ((path << 1) | farBit);
            if (occ[h + nearBit]) stack[++top] = ((u32)(depth + 1) << 16) | ((path << 1) | nearBit);
        }
    }
    // one append per warp
    const unsigned FULL = 0xFFFFFFFFu, lane = threadIdx.x & 31u;
    u32 incl = cnt;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) { u32 v = __shfl_up_sync(FULL, incl, k); if ((int)lane >= k) incl += v; }
    const u32 total = __shfl_sync(FULL, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(count, (unsigned long long)total);
    base = __shfl_sync(FULL, base, 31);
    const unsigned long long at = base + (incl - cnt);
    for (u32 k = 0; k < cnt; k++)
    {
        const unsigned long long pos = at + k;
        if (pos < capacity) { keys[pos] = (k << 16) | list[k]; rayIdx[pos] = (u32)i; }
    }
}

template <int MODE>
__global__ void brick_resolve_kernel(const bmRay* __restrict__ rays, u64 n, const unsigned long long* __restrict__ rayKey, const float4* __restrict__ soup,
                                     bmHit* __restrict__ hits)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = rayKey[i];
    float4 out = make_float4(FLT_MAX, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu));
    if (k != ~0ull)
    {
        const u32 tri = (u32)k;
        const float* r = reinterpret_cast<const float*>(rays + i);
        const V3 o = mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2)), d = mk3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
        const float4 a = __ldg(soup + 3 * (size_t)tri), b = __ldg(soup + 3 * (size_t)tri + 1), c = __ldg(soup + 3 * (size_t)tri + 2);
        const V3 v0 = mk3(a.x, a.y, a.z);
        float u, v;
        const float t = tri_intersect(o, d, v0, mk3(b.x, b.y, b.z) - v0, mk3(c.x, c.y, c.z) - v0, u, v);      // e1, e2 as pack_leaf_tris forms them
        out = make_float4(t, u, v, __uint_as_float(tri));
    }
    reinterpret_cast<float4*>(hits)[i] = out;
}
} // namespace

int brick_table(const TreeView& tv, int bits, u32* d_roots, unsigned char* d_occ, cudaStream_t s)
{
    if (bits < 3 || bits > BRICK_MAX_BITS || bits > tv.grid.D) { set_error("brick bits out of range"); return BM_ERROR_INVALID_PARAMETER; }
    brick_table_kernel<<<div_up((size_t)1 << bits, 128), 128, 0, s>>>(tv, bits, d_roots, d_occ);
    BM_LAUNCHED();
    for (int depth = bits - 1; depth >= 0; --depth)
    {
        brick_occ_kernel<<<div_up((size_t)1 << depth, 128), 128, 0, s>>>(depth, d_occ);
        BM_LAUNCHED();
    }
    return BM_OK;
}

int brick_segments(const TreeView& tv, int bits, const u32* d_roots, const unsigned char* d_occ, const bmRay* d_rays, u64 n, int mode,
                   u32* d_keys, u32* d_rayIdx, u32 capacity, unsigned long long* d_count, cudaStream_t s)
{
    if (!n) return BM_OK;
    BM_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
    float ext = 0.f;
    for (int a = 0; a < 3; a++) ext = fmaxf(ext, tv.grid.wext[a]);
    const float pad = ext * (1.f / 65536.f);
    if (mode == BM_TRACE_REFERENCE) brick_segments_kernel<BM_TRACE_REFERENCE><<<div_up(n, 256), 256, 0, s>>>(tv.grid, bits, d_roots, d_occ, d_rays, n, pad, d_keys, d_rayIdx, capacity, d_count);
    else brick_segments_kernel<BM_TRACE_CLOSEST><<<div_up(n, 256), 256, 0, s>>>(tv.grid, bits, d_roots, d_occ, d_rays, n, pad, d_keys, d_rayIdx, capacity, d_count);
    BM_LAUNCHED();
    return BM_OK;
}

int brick_resolve(const bmRay* d_rays, u64 n, int mode, const unsigned long long* d_rayKey, const float4* d_soup, bmHit* d_hits, cudaStream_t s)
{
    if (!n) return BM_OK;
    if (mode == BM_TRACE_REFERENCE) brick_resolve_kernel<BM_TRACE_REFERENCE><<<div_up(n, 256), 256, 0, s>>>(d_rays, n, d_rayKey, d_soup, d_hits);
    else brick_resolve_kernel<BM_TRACE_CLOSEST><<<div_up(n, 256), 256, 0, s>>>(d_rays, n, d_rayKey, d_soup, d_hits);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
