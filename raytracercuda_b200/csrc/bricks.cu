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
#include <cstdlib>
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

// Bricks and the implicit nodes above them are addressed by their cell coordinates: index(depth; ix, iy, iz) = ix << (by + bz) | iy << bz | iz
// with bx, by, bz = splits on each axis among depths 0..depth-1; occ[(1 << depth) + index] is the node's occupancy, roots[index] at
// depth `bits` the brick's subtree.
__device__ __forceinline__ u32 cell_index(int depth, u32 ix, u32 iy, u32 iz)
{
    const int q = depth / 3, r = depth - 3 * q;
    const int by = q + (r > 1 ? 1 : 0), bz = q;
    return (ix << (by + bz)) | (iy << bz) | iz;
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
    u32 ix, iy, iz; decode_path(bits, b, ix, iy, iz);       // b is the brick's path code (the leading bits of its leaves' codes)
    const u32 at = cell_index(bits, ix, iy, iz);
    roots[at] = ref;
    occ[(1u << bits) + at] = ref != BRICK_EMPTY;
}

// occupancy of the implicit levels above the bricks, one level per launch (a level has at most 2^14 nodes)
__global__ void brick_occ_kernel(int depth, unsigned char* occ)
{
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;      // path code of the depth-`depth` node
    if (p >= (1u << depth)) return;
    u32 c[3]; decode_path(depth, p, c[0], c[1], c[2]);
    const int axis = depth % 3;
    u32 lo[3] = { c[0], c[1], c[2] }, hi[3] = { c[0], c[1], c[2] };
    lo[axis] = 2u * c[axis]; hi[axis] = 2u * c[axis] + 1u;
    occ[(1u << depth) + cell_index(depth, c[0], c[1], c[2])] = occ[(2u << depth) + cell_index(depth + 1, lo[0], lo[1], lo[2])] | occ[(2u << depth) + cell_index(depth + 1, hi[0], hi[1], hi[2])];
}

// The implicit top of the reference's traversal for one ray (oracle.cpp march_ray / BuildTree.cu:394-481 restricted to depth <= B):
// pop a node, test its cell with the reference's slab arithmetic, order the two children by the side of the split plane on which the ray
// enters the cell, push far then near; a node at depth B is a brick and is noted instead.  Children without any leaf are not pushed
// (the reference has no such node).  CLOSEST: the direct traversal culls with the boxes of whole triangles, which reach beyond their
// cells by at most a triangle; a hit point itself always lies in a cell of its triangle, up to rounding, so here the cells are grown by
// `pad` and every brick the forward ray touches is noted (the order only affects how early the search interval shrinks).
// Segments are traced in WAVES of their position in the ray's order (0-1, 2-4, 5-9, 10+), brick by brick inside a wave: what a ray found in
// an earlier wave is known when its later segments come up, so they are skipped (REFERENCE: the first brick with a hit decides) or start
// with the search interval already cut (CLOSEST).  Without waves a batch with a 69 % hit rate walked 46 % more nodes than the direct traversal.
__device__ __forceinline__ u32 seg_wave(u32 k) { return k < 2u ? 0u : (k < 5u ? 1u : (k < 10u ? 2u : 3u)); }

template <int MODE>
__global__ void __launch_bounds__(256) brick_segments_kernel(Grid g, int bits, const u32* __restrict__ roots, const unsigned char* __restrict__ occ,
                                                              const bmRay* __restrict__ rays, u64 n, float pad, int cohBits, V3 cohScale, u32* __restrict__ keys, u32* __restrict__ rayIdx,
                                                              u32 capacity, unsigned long long* count)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 list[BR_MAX_SEGS];                                 // brick index | coherence key << 16
    u32 cnt = 0;
    if (i < n)
    {
        const float* r = reinterpret_cast<const float*>(rays + i);
        const V3 o = mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2)), d = mk3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
        const V3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
        // The cell of a child shares five of its six slab planes with its parent, so the slab distances (bMin - o) * inv, (bMax - o) * inv
        // of bmBoxRayIntersect (CudaComon.cuh:158-172) are carried down and only the split plane's is computed per level -- the same
        // operations on the same operands as evaluating the child's box from scratch, hence the same bits.  The near child continues in
        // registers; the far one waits on a small stack with its slab distances.
        u32 sE[BRICK_MAX_BITS + 1];
        float sT[BRICK_MAX_BITS + 1][6];
        int top = -1;
        u32 e = 0u;                                       // depth << 24 | ix << 16 | iy << 8 | iz
        float tlo[3], thi[3];                             // (bMin - o) * inv, (bMax - o) * inv per axis
        {
            const float oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
#pragma unroll
            for (int a = 0; a < 3; a++)
            {
                float lo = grid_at(g, a, 0u), hi = grid_at(g, a, 1u << g.k[a]);
                if (MODE != BM_TRACE_REFERENCE) { lo -= pad; hi += pad; }
                tlo[a] = (lo - oo[a]) * ii[a]; thi[a] = (hi - oo[a]) * ii[a];
            }
        }
        for (;;)
        {
            bool pop = true;
            // bmBoxRayIntersect on the carried slab distances
            const float ftmax = fmin_(fmax_(tlo[0], thi[0]), fmin_(fmax_(tlo[1], thi[1]), fmax_(tlo[2], thi[2])));
            if (!(ftmax < 0.f))
            {
                const float ftmin = fmax_(fmin_(tlo[0], thi[0]), fmax_(fmin_(tlo[1], thi[1]), fmin_(tlo[2], thi[2])));
                if (ftmax >= ftmin)
                {
                    const float boxDist = fmax_(0.f, ftmin);
                    const int depth = (int)(e >> 24);
                    const u32 ix = (e >> 16) & 0xFFu, iy = (e >> 8) & 0xFFu, iz = e & 0xFFu;
                    const int q = depth / 3, axis = depth - 3 * q;
                    const int by = q + (axis > 1 ? 1 : 0), bz = q;
                    if (depth == bits)
                    {
                        // coherence key: where the ray enters the brick (3 bits per axis below the brick's own coordinate bits, interleaved)
                        // and its direction octant -- segments of one brick are sorted by it, so that the lanes of a warp walk neighbouring
                        // rays through the same nodes.  Only an ordering hint: plain arithmetic, no exactness needed.
                        u32 coh = 0u;
                        if (cohBits)
                        {
                            const int bx = q + (axis > 0 ? 1 : 0);
                            const float ex = o.x + boxDist * d.x, ey = o.y + boxDist * d.y, ez = o.z + boxDist * d.z;
                            const u32 lx = (u32)fminf(fmaxf((ex - g.wmin[0]) * cohScale.x, 0.f), 2047.f), ly = (u32)fminf(fmaxf((ey - g.wmin[1]) * cohScale.y, 0.f), 2047.f),
                                      lz = (u32)fminf(fmaxf((ez - g.wmin[2]) * cohScale.z, 0.f), 2047.f);
                            const u32 sx3 = (lx >> max(g.k[0] - bx - 3, 0)) & 7u, sy3 = (ly >> max(g.k[1] - by - 3, 0)) & 7u, sz3 = (lz >> max(g.k[2] - bz - 3, 0)) & 7u;
                            const u32 oct = (d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u);
                            coh = ((oct << 9) | (spread3(sx3) << 2) | (spread3(sy3) << 1) | spread3(sz3)) >> (12 - cohBits);
                        }
                        if (cnt < (u32)BR_MAX_SEGS) list[cnt++] = ((ix << (by + bz)) | (iy << bz) | iz) | (coh << 16);
                    }
                    else
                    {
                        const u32 ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
                        const int sh = (axis == 0 ? g.k[0] : (axis == 1 ? g.k[1] : g.k[2])) - q - 1;      // leaf-level shift of the children's cells on this axis
                        const u32 bi = ((ia << 1) | 1u) << sh;
                        const float split = axis == 0 ? grid_at(g, 0, bi) : (axis == 1 ? grid_at(g, 1, bi) : grid_at(g, 2, bi));      // = .5f * (max + min) of the cell, BuildTree.cu:443
                        const float oa = get(o, axis), ia_ = get(inv, axis);
                        const float p = oa + boxDist * get(d, axis);
                        const bool lowFirst = p < split;
                        // the split plane as the low child's bMax and the high child's bMin (CLOSEST: each grown by the pad)
                        const float tLowMax = ((MODE != BM_TRACE_REFERENCE ? split + pad : split) - oa) * ia_;
                        const float tHighMin = ((MODE != BM_TRACE_REFERENCE ? split - pad : split) - oa) * ia_;
                        const u32 cx = axis == 0 ? 2u * ix : ix, cy = axis == 1 ? 2u * iy : iy, cz = axis == 2 ? 2u * iz : iz;
                        const int cby = by + (axis == 1 ? 1 : 0), cbz = bz + (axis == 2 ? 1 : 0);
                        const u32 lowIdx = (cx << (cby + cbz)) | (cy << cbz) | cz;
                        const u32 stepIdx = axis == 0 ? (1u << (cby + cbz)) : (axis == 1 ? (1u << cbz) : 1u);      // index distance between the two children
                        const u32 lowEntry = ((u32)(depth + 1) << 24) | (cx << 16) | (cy << 8) | cz;
                        const u32 stepEntry = axis == 0 ? (1u << 16) : (axis == 1 ? (1u << 8) : 1u);
                        const unsigned char* oc = occ + (2u << depth) + lowIdx;
                        const bool hasLow = oc[0] != 0, hasHigh = oc[stepIdx] != 0;
                        const bool nearLow = lowFirst ? hasLow || !hasHigh : !hasHigh;      // the child to continue with: the near one if it exists, else the other
                        if (hasLow && hasHigh)
                        {
                            // far child onto the stack with its slab distances
                            ++top;
                            sE[top] = lowFirst ? lowEntry + stepEntry : lowEntry;
#pragma unroll
                            for (int a = 0; a < 3; a++) { sT[top][a] = tlo[a]; sT[top][3 + a] = thi[a]; }
                            if (lowFirst) sT[top][axis] = tHighMin; else sT[top][3 + axis] = tLowMax;
                        }
                        if (hasLow || hasHigh)
                        {
                            pop = false;
                            if (nearLow) { e = lowEntry; if (axis == 0) thi[0] = tLowMax; else if (axis == 1) thi[1] = tLowMax; else thi[2] = tLowMax; }
                            else { e = lowEntry + stepEntry; if (axis == 0) tlo[0] = tHighMin; else if (axis == 1) tlo[1] = tHighMin; else tlo[2] = tHighMin; }
                        }
                    }
                }
            }
            if (pop)
            {
                if (top < 0) break;
                e = sE[top];
#pragma unroll
                for (int a = 0; a < 3; a++) { tlo[a] = sT[top][a]; thi[a] = sT[top][3 + a]; }
                --top;
            }
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
        if (pos < capacity) { keys[pos] = (k << 24) | (((seg_wave(k) << bits) | (list[k] & 0xFFFFu)) << cohBits) | (list[k] >> 16); rayIdx[pos] = (u32)i; }
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

int brick_coherence_bits(int bits)
{
    // 0 (default) = the segments of a brick stay in ray order.  Measured on the 20 M-triangle soup, 8 M incoherent rays: 12 bits of
    // entry-point / octant order change nothing (116.4 vs 116.9 Mrays/s reference, 158.7 vs 158.9 closest; profiles/r2_SUMMARY.md): the
    // walk is bound by instruction issue per node visit, not by how many lanes share a node.
    static const int want = getenv("BM_BRICK_COHERENCE") ? atoi(getenv("BM_BRICK_COHERENCE")) : 0;
    int c = want < 0 ? 0 : (want > 12 ? 12 : want);
    if (c > 24 - 2 - bits) c = 24 - 2 - bits;
    return c < 0 ? 0 : c;
}

int brick_segments(const TreeView& tv, int bits, const u32* d_roots, const unsigned char* d_occ, const bmRay* d_rays, u64 n, int mode,
                   u32* d_keys, u32* d_rayIdx, u32 capacity, unsigned long long* d_count, cudaStream_t s)
{
    if (!n) return BM_OK;
    BM_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
    float ext = 0.f;
    for (int a = 0; a < 3; a++) ext = fmaxf(ext, tv.grid.wext[a]);
    const float pad = ext * (1.f / 65536.f);
    const int cohBits = brick_coherence_bits(bits);
    const V3 cohScale = mk3((float)(1u << tv.grid.k[0]) / tv.grid.wext[0], (float)(1u << tv.grid.k[1]) / tv.grid.wext[1], (float)(1u << tv.grid.k[2]) / tv.grid.wext[2]);
    if (mode == BM_TRACE_REFERENCE) brick_segments_kernel<BM_TRACE_REFERENCE><<<div_up(n, 256), 256, 0, s>>>(tv.grid, bits, d_roots, d_occ, d_rays, n, pad, cohBits, cohScale, d_keys, d_rayIdx, capacity, d_count);
    else brick_segments_kernel<BM_TRACE_CLOSEST><<<div_up(n, 256), 256, 0, s>>>(tv.grid, bits, d_roots, d_occ, d_rays, n, pad, cohBits, cohScale, d_keys, d_rayIdx, capacity, d_count);
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
