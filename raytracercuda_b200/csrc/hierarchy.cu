// hierarchy.cu -- from sorted (key, triangle) pairs to the traversable tree:
//   leaf segmentation (distinct codes)  ->  binary radix tree over the distinct codes, built bottom-up in one pass together with
//   the tight boxes (agglomerate_kernel)  ->  128-byte nodes.
//
// Topology equivalence with the reference: the reference kd node at depth d with TWO children is
// exactly the radix-tree node whose keys share a d-bit prefix and then differ; single-child chain nodes of the reference
// have no counterpart because they take no decision.  agglomerate_kernel stores (prefix, prefixLen) and, precomputed from the
// grid tables, the node's cell box, split axis (prefixLen % 3) and split plane -- the reference's own values.
#include <cstdlib>
#include "bm_device.cuh"
#include "bm_wide.cuh"

namespace bm
{
namespace
{
__global__ void mark_heads_kernel(const u64* __restrict__ keys, size_t n, int subBits, u32* __restrict__ flags)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || (keys[i] >> subBits) != (keys[i - 1] >> subBits)) ? 1u : 0u;
}

__global__ void write_leaves_kernel(Grid g, const u64* __restrict__ keys, const u32* __restrict__ flags, const u32* __restrict__ leafIdx, size_t n,
                                    u32 numLeaves, float4* __restrict__ leaves, u32* __restrict__ clusterCounts, u32* __restrict__ codes)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!flags[i]) return;
    // run length: scan forward to the next head (leaves are short: <= a few hundred pairs)
    size_t j = i + 1;
    while (j < n && !flags[j]) ++j;
    u32 l = leafIdx[i], code = (u32)(keys[i] >> g.subBits), cnt = (u32)(j - i);
    u32 ix, iy, iz; decode_path(g.D, code, ix, iy, iz);
    V3 mn, mx; cell_box(g, g.D, ix, iy, iz, mn, mx);
    float4* o = leaves + LEAF_F4 * (size_t)l;
    o[0] = make_float4(__uint_as_float(code), __uint_as_float((u32)i), __uint_as_float(cnt), __uint_as_float(0u));
    o[1] = make_float4(mn.x, mn.y, mn.z, mx.x);
    o[2] = make_float4(mx.y, mx.z, 0.f, 0.f);
    clusterCounts[l] = (cnt + CLUSTER - 1) / CLUSTER;
    codes[l] = code;                                   // dense copy for the hierarchy's binary searches
}

__global__ void pack_leaf_tris_kernel(const float4* __restrict__ soup, const u32* __restrict__ vals, size_t n, float4* __restrict__ ltris)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* p = soup + 3 * (size_t)vals[i];
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    float4* o = ltris + 3 * i;
    // e1 = v1 - v0, e2 = v2 - v0: the first two operations of bmTriIntersect (CudaComon.cuh:122-123)
    o[0] = a;
    o[1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, b.w);
    o[2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, c.w);
}

// Cluster boxes: one thread per pair; the thread of the first pair of each group of CLUSTER consecutive pairs of a leaf builds the
// group's box.  Boxes enclose WHOLE triangles (not clipped to the cell: reference mode must still find a hit that lies outside the
// cell, BuildTree.cu:413-431).  v1 = v0 + e1 may differ from the stored vertex by one rounding, and a hit the triangle test accepts
// numerically may lie a few ulps of the ray length outside the exact box: `pad` (world extent * 2^-20, far below the cell size)
// absorbs both.
__global__ void cluster_bounds_kernel(const float4* __restrict__ leaves, const u32* __restrict__ flags, const u32* __restrict__ leafIdx,
                                      const u32* __restrict__ clusterFirst, size_t n, const float4* __restrict__ ltris, float4* __restrict__ clusters, float pad)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 l = leafIdx[i] + flags[i] - 1u;        // leafIdx = exclusive scan of the head flags: heads before i, so a non-head pair sees its leaf + 1
    float4 r0 = __ldg(&leaves[LEAF_F4 * (size_t)l]);
    u32 first = __float_as_uint(r0.y), cnt = __float_as_uint(r0.z);
    u32 local = (u32)(i - first);
    if (local % CLUSTER) return;
    float clo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, chi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    u32 end = min(local + CLUSTER, cnt);
    for (u32 k = local; k < end; k++)
    {
        const float4* p = ltris + 3 * (size_t)(first + k);
        float4 a = __ldg(p), e1 = __ldg(p + 1), e2 = __ldg(p + 2);
        float v[3][3] = { { a.x, a.y, a.z }, { a.x + e1.x, a.y + e1.y, a.z + e1.z }, { a.x + e2.x, a.y + e2.y, a.z + e2.z } };
#pragma unroll
        for (int q = 0; q < 3; q++)
#pragma unroll
            for (int c = 0; c < 3; c++) { clo[c] = fminf(clo[c], v[q][c] - pad); chi[c] = fmaxf(chi[c], v[q][c] + pad); }
    }
    float4* o = clusters + CLUSTER_F4 * (size_t)(clusterFirst[l] + local / CLUSTER);
    o[0] = make_float4(clo[0], clo[1], clo[2], chi[0]);
    o[1] = make_float4(chi[1], chi[2], 0.f, 0.f);
}

// Per leaf: tight box = union of its cluster boxes, first-cluster index into the record, leaf-size statistics (stats[0] = max leaf
// size, stats[1] = leaves above 256 faces) and the reference's face cap.
__global__ void leaf_bounds_kernel(float4* __restrict__ leaves, u32 numLeaves, const u32* __restrict__ clusterFirst, float4* __restrict__ ltris,
                                   const float4* __restrict__ clusters, float4* __restrict__ box2, u32* stats, u32* __restrict__ overList)
{
    u32 l = blockIdx.x * blockDim.x + threadIdx.x;
    float4 r0 = l < numLeaves ? leaves[LEAF_F4 * (size_t)l] : make_float4(0, 0, 0, 0);
    u32 first = __float_as_uint(r0.y), cnt = l < numLeaves ? __float_as_uint(r0.z) : 0u;
    u32 wmax = __reduce_max_sync(0xFFFFFFFFu, cnt);
    u32 wover = __popc(__ballot_sync(0xFFFFFFFFu, cnt > 256u));
    if ((threadIdx.x & 31) == 0)
    {
        if (wmax > __ldcg(&stats[0])) atomicMax(&stats[0], wmax);
        if (wover) atomicAdd(&stats[1], wover);
    }
    if (l >= numLeaves) return;
    u32 cf = clusterFirst[l];
    leaves[LEAF_F4 * (size_t)l].w = __uint_as_float(cf);
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    u32 ncl = (cnt + CLUSTER - 1) / CLUSTER;
    for (u32 c = 0; c < ncl; c++)
    {
        const float4* p = clusters + CLUSTER_F4 * (size_t)(cf + c);
        float4 b0 = p[0], b1 = p[1];
        lo[0] = fminf(lo[0], b0.x); lo[1] = fminf(lo[1], b0.y); lo[2] = fminf(lo[2], b0.z);
        hi[0] = fmaxf(hi[0], b0.w); hi[1] = fmaxf(hi[1], b1.x); hi[2] = fmaxf(hi[2], b1.y);
    }
    box2[2 * (size_t)l] = make_float4(lo[0], lo[1], lo[2], hi[0]);
    box2[2 * (size_t)l + 1] = make_float4(hi[1], hi[2], 0.f, 0.f);
    // face cap: leaves above 256 faces are queued for cap_flags_kernel
    if (cnt > 256u) overList[atomicAdd(&stats[2], 1u)] = l;
}

// The reference keeps the first 256 faces inserted into a leaf (BuildTree.cu:53-60) = the 256 smallest (mesh, triangle) ids in its
// serial mode; faces beyond that are never tested.  One block per queued leaf: rank of every face id among the leaf's ids, faces of rank
// >= 256 get the EXCLUDED bit.  (O(n^2 / 256) per thread; such leaves only exist for meshes far denser than the reference can hold.)
constexpr int CAP_THREADS = 256;
constexpr u32 CAP_SMEM_IDS = 8192;
__global__ void __launch_bounds__(CAP_THREADS) cap_flags_kernel(const float4* __restrict__ leaves, const u32* __restrict__ overList, const u32* stats,
                                                                float4* __restrict__ ltris)
{
    __shared__ u32 ids[CAP_SMEM_IDS];
    const u32 nOver = stats[2];
    for (u32 item = blockIdx.x; item < nOver; item += gridDim.x)
    {
        float4 r0 = leaves[LEAF_F4 * (size_t)overList[item]];
        u32 first = __float_as_uint(r0.y), cnt = __float_as_uint(r0.z);
        u32 inSmem = min(cnt, CAP_SMEM_IDS);
        __syncthreads();
        for (u32 i = threadIdx.x; i < inSmem; i += CAP_THREADS) ids[i] = __float_as_uint(ltris[3 * (size_t)(first + i)].w);
        __syncthreads();
        for (u32 i = threadIdx.x; i < cnt; i += CAP_THREADS)
        {
            u32 ti = i < inSmem ? ids[i] : __float_as_uint(ltris[3 * (size_t)(first + i)].w), rank = 0;
            for (u32 j = 0; j < inSmem && rank < 256u; j++) rank += ids[j] < ti ? 1u : 0u;
            for (u32 j = inSmem; j < cnt && rank < 256u; j++) rank += __float_as_uint(ltris[3 * (size_t)(first + j)].w) < ti ? 1u : 0u;
            if (rank >= 256u)
            {
                float4* p = ltris + 3 * (size_t)(first + i) + 1;
                p->w = __uint_as_float(__float_as_uint(p->w) | EXCLUDED_BIT);
            }
        }
    }
}

// ---- hierarchy emission + refit in one bottom-up pass ---------------------------------------------------------------------------
// Apetrei 2014 ("Fast and simple agglomerative LBVH construction"): the binary radix tree over the sorted distinct codes, built from
// the leaves upwards.  A thread owning the key range [l, r] (initially one leaf) joins the neighbour range across the boundary with
// the smaller delta: delta(i) = code[i] ^ code[i+1]; if delta(r) < delta(l-1) its parent is internal node r (it is the LEFT child and
// the parent's split lies between r and r+1), else node l-1 (RIGHT child).  One atomic exchange per node on the "other bound" tells
// the second arrival (which then owns the merged range and both child boxes) from the first (which stops).  This is the same tree
// Karras' top-down searches produce (node = split position), without the 2 log L random reads per node, and the tight boxes are
// united on the way up, so no separate refit pass is needed.  Each node also receives the reference node it stands for: common
// prefix (prefixLen bits of the codes of its range), cell box and split plane from the grid (BuildTree.cu:436-443).
__device__ __forceinline__ u32 range_delta(const u32* __restrict__ codes, int i, int n)      // boundary between leaf i and i+1
{
    return (i < 0 || i >= n - 1) ? 0xFFFFFFFFu : (__ldg(codes + i) ^ __ldg(codes + i + 1));
}

struct Box2 { float4 a, b; };                 // (lo.xyz, hi.x) (hi.y, hi.z, -, -)
__device__ __forceinline__ Box2 load_box(const float4* p) { Box2 r; r.a = __ldcg(p); r.b = __ldcg(p + 1); return r; }   // written by another thread of this launch: bypass L1

__device__ __forceinline__ WSlotBox slot_box(const Box2& b)
{
    WSlotBox r;
    r.lo[0] = b.a.x; r.lo[1] = b.a.y; r.lo[2] = b.a.z; r.hi[0] = b.a.w; r.hi[1] = b.b.x; r.hi[2] = b.b.y;
    return r;
}

// What a finished node receives (both phases of agglomerate_kernel): the 128-byte binary record (nodes, nullable) and nodeMeta[n] =
// (packed cell coordinates, depth).  The 64-byte wide records are written by wide_encode_kernel below from childRefs / nodeMeta / the
// boxes this pass leaves behind.
__device__ __forceinline__ void finish_node(const Grid& g, int D, const u32* __restrict__ codes, int parent, int l, int r, u32 c0, u32 c1, const Box2& b0, const Box2& b1,
                                            float4* nodes, uint2* nodeMeta)
{
    u32 cl = __ldg(codes + l), cr = __ldg(codes + r);
    int plen = __clz(cl ^ cr) - (32 - D);         // common prefix length inside the D-bit code (l < r: codes differ)
    u32 prefix = plen > 0 ? (cl >> (D - plen)) : 0u;
    u32 ix, iy, iz; decode_path(plen, prefix, ix, iy, iz);
    if (nodes)
    {
        V3 mn, mx; cell_box(g, plen, ix, iy, iz, mn, mx);
        int axis = plen % 3;
        u32 ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
        int sh = g.k[axis] - axis_bits(plen, axis);   // >= 1: the node splits on this axis
        float sp = grid_at(g, axis, ((ia << 1) | 1u) << (sh - 1));
        float4* np = nodes + NODE_F4 * (size_t)parent;
        np[0] = b0.a;                                                   // lo0.xyz hi0.x
        np[1] = make_float4(b0.b.x, b0.b.y, b1.a.x, b1.a.y);            // hi0.yz lo1.xy
        np[2] = make_float4(b1.a.z, b1.a.w, b1.b.x, b1.b.y);            // lo1.z hi1.xyz
        np[3] = make_float4(__uint_as_float(c0), __uint_as_float(c1), __uint_as_float(prefix), __uint_as_float((u32)plen));
        np[4] = make_float4(mn.x, mn.y, mn.z, sp);
        np[5] = make_float4(mx.x, mx.y, mx.z, __uint_as_float((u32)axis));
    }
    // (cell coordinates, depth) of the node: what the wide-record pass and bmTreeCopyTopology read
    __stcg(&nodeMeta[parent], make_uint2(wide_pack_cell(g, ix, iy, iz), (u32)plen));
}
__device__ __forceinline__ Box2 unite(const Box2& b0, const Box2& b1)
{
    Box2 m;
    m.a = make_float4(fminf(b0.a.x, b1.a.x), fminf(b0.a.y, b1.a.y), fminf(b0.a.z, b1.a.z), fmaxf(b0.a.w, b1.a.w));
    m.b = make_float4(fmaxf(b0.b.x, b1.b.x), fmaxf(b0.b.y, b1.b.y), 0.f, 0.f);
    return m;
}

// One 64-bit exchange per arrival carries everything the two children of a node have to tell each other: (subtree reference, far bound of
// its key range).  The second arrival therefore learns its sibling from the exchange itself -- its loads of the sibling's box depend on
// the value returned, so they cannot be issued early and need no fence -- and is the one that writes the node's child references.  A
// device-wide fence remains only where a thread has written something another thread will read: the box of a NODE it has just finished,
// before it arrives at that node's parent; a leaf's box was there before the launch.  That is one fence per internal node instead of
// three (the pass was bound by them: membar 11.5 of 24 stall cycles per issue, profiles/r2_SUMMARY.md).
// Tried and dropped: meeting through shared memory for nodes whose range lies inside the block's leaves (no fences at all for ~94 % of
// the nodes).  Finding out whether a parent is such a node needs the far end of the sibling's range -- a walk over the deltas as long as
// the sibling is large, by the few lanes still climbing: 2.7x the instructions, 10.5 instead of 8.3 ms on 50 M leaves.
constexpr int AGG_THREADS = 128;
__global__ void __launch_bounds__(AGG_THREADS) agglomerate_kernel(Grid g, const u32* __restrict__ codes, u32 numLeaves, const float4* __restrict__ leafBox, float4* nodeBox,
                                                                  unsigned long long* meet, u32* childRefs, float4* nodes, uint2* nodeMeta, u32* rootOut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = (int)numLeaves;
    if (i >= n) return;
    const int D = g.D;
    int l = i, r = i;
    u32 cur = (u32)i | LEAF_BIT;
    Box2 box; box.a = __ldg(leafBox + 2 * (size_t)i); box.b = __ldg(leafBox + 2 * (size_t)i + 1);
    for (;;)
    {
        const bool leftChild = (l == 0) || (r != n - 1 && range_delta(codes, r, n) < range_delta(codes, l - 1, n));
        const int parent = leftChild ? r : l - 1;
        // release: nodeBox[cur], written below in the previous round, is visible before anybody can learn of `cur`.  A release fence is enough
        // for that (__threadfence() is the sequentially consistent one: MEMBAR.SC.GPU + CCTL.IVALL; this is MEMBAR.ALL.GPU, no L1 invalidate)
        if (!(cur & LEAF_BIT)) asm volatile("fence.release.gpu;" ::: "memory");
        const unsigned long long other = atomicExch(&meet[parent], ((unsigned long long)cur << 32) | (u32)(leftChild ? l : r));
        if (other == ~0ull) return;                   // first arrival: the sibling will finish this node
        const u32 sib = (u32)(other >> 32);
        if (leftChild) r = (int)(u32)other; else l = (int)(u32)other;
        const Box2 sb = load_box((sib & LEAF_BIT) ? leafBox + 2 * (size_t)(sib & ~LEAF_BIT) : nodeBox + 2 * (size_t)sib);
        const u32 c0 = leftChild ? cur : sib, c1 = leftChild ? sib : cur;
        const Box2& b0 = leftChild ? box : sb; const Box2& b1 = leftChild ? sb : box;
        __stcg(reinterpret_cast<uint2*>(childRefs) + parent, make_uint2(c0, c1));
        finish_node(g, D, codes, parent, l, r, c0, c1, b0, b1, nodes, nodeMeta);
        if (l == 0 && r == n - 1) { *rootOut = (u32)parent; return; }
        box = unite(b0, b1);
        float4* o = nodeBox + 2 * (size_t)parent;       // read by the thread that finishes the parent, and by wide_encode_kernel
        __stcg(o, box.a); __stcg(o + 1, box.b);
        cur = (u32)parent;
    }
}

// The 64-byte wide records (bm_wide.cuh), one thread per internal node AFTER the bottom-up pass: node b's record absorbs its internal
// children (their own children become b's slots), so it needs the child references, cell coordinates and boxes of two levels below b.
// Inside the climb only the second arrival of every node is alive, a handful of lanes per warp, and the encoder (a search for the
// quantisation grid, ~400 instructions) ran at that occupancy: 151 of 218 ms of a 100 M-triangle build.  Here every lane has a node.
__global__ void __launch_bounds__(128) wide_encode_kernel(Grid g, u32 numNodes, const u32* __restrict__ childRefs, const uint2* __restrict__ nodeMeta,
                                                          const float4* __restrict__ leafBox, const float4* __restrict__ nodeBox, uint4* __restrict__ wnodes)
{
    const u32 n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= numNodes) return;
    const uint2 m0 = __ldg(&nodeMeta[n]);
    const uint2 cc = __ldg(reinterpret_cast<const uint2*>(childRefs) + n);
    const u32 c[2] = { cc.x, cc.y };
    u32 ref[4] = { c[0], WREF_EMPTY, c[1], WREF_EMPTY };
    WSlotBox box[4];
    u32 cellAB[2] = { 0u, 0u }, lenAB[2] = { WLEN_NONE, WLEN_NONE };
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        if (c[h] & LEAF_BIT)
        {
            const float4* p = leafBox + 2 * (size_t)(c[h] & ~LEAF_BIT);
            Box2 b; b.a = __ldg(p); b.b = __ldg(p + 1);
            box[2 * h] = slot_box(b); box[2 * h + 1] = box[2 * h];
            continue;
        }
        const uint2 m = __ldg(&nodeMeta[c[h]]);
        cellAB[h] = m.x; lenAB[h] = m.y;
        const uint2 gg = __ldg(reinterpret_cast<const uint2*>(childRefs) + c[h]);
        ref[2 * h] = gg.x; ref[2 * h + 1] = gg.y;
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
            const u32 gref = k ? gg.y : gg.x;
            const float4* p = (gref & LEAF_BIT) ? leafBox + 2 * (size_t)(gref & ~LEAF_BIT) : nodeBox + 2 * (size_t)gref;
            Box2 b; b.a = __ldg(p); b.b = __ldg(p + 1);
            box[2 * h + k] = slot_box(b);
        }
    }
    WQ rec[4];
    wide_encode(g, m0.x, m0.y, cellAB[0], lenAB[0], cellAB[1], lenAB[1], ref, box, rec);
    uint4* wp = wnodes + WNODE_U4 * (size_t)n;
#pragma unroll
    for (int k = 0; k < 4; k++) wp[k] = make_uint4(rec[k].x, rec[k].y, rec[k].z, rec[k].w);
}

__global__ void copy_topology_kernel(Grid g, const u32* __restrict__ childRefs, const uint2* __restrict__ meta, u32 n, u32* __restrict__ plen, u32* __restrict__ prefix,
                                     u32* __restrict__ c0, u32* __restrict__ c1)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 m = meta[i];
    if (plen) plen[i] = m.y;
    if (prefix) { u32 ix, iy, iz; wide_unpack_cell(g, m.x, ix, iy, iz); prefix[i] = m.y ? encode_path((int)m.y, ix, iy, iz) : 0u; }
    if (c0) c0[i] = childRefs[2 * (size_t)i];
    if (c1) c1[i] = childRefs[2 * (size_t)i + 1];
}

__global__ void copy_child_boxes_kernel(const u32* __restrict__ childRefs, const float4* __restrict__ leafBox, const float4* __restrict__ nodeBox, u32 n, float* __restrict__ out)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        u32 c = childRefs[2 * (size_t)i + h];
        const float4* p = (c & LEAF_BIT) ? leafBox + 2 * (size_t)(c & ~LEAF_BIT) : nodeBox + 2 * (size_t)c;
        float4 a = p[0], b = p[1];
        float* o = out + 12 * (size_t)i + 6 * h;
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y;
    }
}

__global__ void copy_wide_slots_kernel(Grid g, const uint4* __restrict__ wnodes, u32 n, float* __restrict__ boxes, u32* __restrict__ refs)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = wnodes + WNODE_U4 * (size_t)i;
    uint4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    WCell c0 = wide_cell(g, q0.x, (int)(q0.w & 0xFFu));
    WFrame f = wide_frame(g, c0.mn, c0.mx, q3.z, q3.w, q0.w >> 24);
    const u32 LW[3] = { q2.x, q2.y, q2.z }, HW[3] = { q2.w, q3.x, q3.y };
    if (refs) { refs[4 * (size_t)i] = q1.x; refs[4 * (size_t)i + 1] = q1.y; refs[4 * (size_t)i + 2] = q1.z; refs[4 * (size_t)i + 3] = q1.w; }
    if (boxes)
        for (int s = 0; s < 4; s++)
            for (int a = 0; a < 3; a++)
            {
                boxes[24 * (size_t)i + 6 * s + a] = wide_fma(wide_magic(LW[a], s), f.step[a], f.addLo[a]);
                boxes[24 * (size_t)i + 6 * s + 3 + a] = wide_fma(wide_magic(HW[a], s), f.step[a], f.addHi[a]);
            }
}

// (code << subBits | sub, tri)  ->  (code << 32 | tri): the key whose order is the canonical (code, triangle) order
__global__ void canonical_keys_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals, size_t n, int subBits, u64* __restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((keys[i] >> subBits) << 32) | vals[i];
}
__global__ void split_canonical_kernel(const u64* __restrict__ keys, size_t n, u32* __restrict__ code, u32* __restrict__ tri)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (code) code[i] = (u32)(keys[i] >> 32);
    if (tri) tri[i] = (u32)keys[i];
}
} // namespace

int mark_heads(const u64* d_keys, size_t n, int subBits, u32* d_flags, cudaStream_t s)
{
    if (!n) return BM_OK;
    mark_heads_kernel<<<div_up(n, 256), 256, 0, s>>>(d_keys, n, subBits, d_flags);
    BM_LAUNCHED();
    return BM_OK;
}
int write_leaves(const Grid& g, const u64* d_keys, const u32* d_flags, const u32* d_leafIdx, size_t n, u32 numLeaves, float4* d_leaves, u32* d_clusterCounts,
                 u32* d_codes, cudaStream_t s)
{
    if (!n) return BM_OK;
    write_leaves_kernel<<<div_up(n, 256), 256, 0, s>>>(g, d_keys, d_flags, d_leafIdx, n, numLeaves, d_leaves, d_clusterCounts, d_codes);
    BM_LAUNCHED();
    return BM_OK;
}
int pack_leaf_tris(const float4* d_soup, const u32* d_vals, size_t n, float4* d_ltris, cudaStream_t s)
{
    if (!n) return BM_OK;
    pack_leaf_tris_kernel<<<div_up(n, 256), 256, 0, s>>>(d_soup, d_vals, n, d_ltris);
    BM_LAUNCHED();
    return BM_OK;
}
int leaf_bounds(float4* d_leaves, u32 numLeaves, const u32* d_flags, const u32* d_leafIdx, size_t numPairs, const u32* d_clusterFirst, float4* d_ltris,
                float4* d_clusters, float4* d_leafBox2, float pad, u32* d_stats, u32* d_overList, cudaStream_t s)
{
    if (!numLeaves) return BM_OK;
    cluster_bounds_kernel<<<div_up(numPairs, 256), 256, 0, s>>>(d_leaves, d_flags, d_leafIdx, d_clusterFirst, numPairs, d_ltris, d_clusters, pad);
    BM_LAUNCHED();
    leaf_bounds_kernel<<<div_up(numLeaves, 128), 128, 0, s>>>(d_leaves, numLeaves, d_clusterFirst, d_ltris, d_clusters, d_leafBox2, d_stats, d_overList);
    BM_LAUNCHED();
    cap_flags_kernel<<<(unsigned)sm_count() * 4, CAP_THREADS, 0, s>>>(d_leaves, d_overList, d_stats, d_ltris);     // exits at once when nothing is queued
    BM_LAUNCHED();
    return BM_OK;
}
int build_hierarchy(const Grid& g, const u32* d_codes, u32 numLeaves, const float4* d_leafBox2, float4* d_nodeBox2, unsigned long long* d_meet, u32* d_childRefs,
                    float4* d_nodes, uint4* d_wnodes, uint2* d_nodeMeta, u32* d_root, cudaStream_t s)
{
    if (numLeaves < 2) return BM_OK;
    BM_CUDA(cudaMemsetAsync(d_meet, 0xFF, (size_t)(numLeaves - 1) * 8, s));
    agglomerate_kernel<<<div_up(numLeaves, AGG_THREADS), AGG_THREADS, 0, s>>>(g, d_codes, numLeaves, d_leafBox2, d_nodeBox2, d_meet, d_childRefs, d_nodes, d_nodeMeta, d_root);
    BM_LAUNCHED();
    if (d_wnodes)
    {
        wide_encode_kernel<<<div_up(numLeaves - 1, 128), 128, 0, s>>>(g, numLeaves - 1, d_childRefs, d_nodeMeta, d_leafBox2, d_nodeBox2, d_wnodes);
        BM_LAUNCHED();
    }
    return BM_OK;
}
int copy_topology(const Grid& g, const u32* d_childRefs, const uint2* d_nodeMeta, u32 numNodes, u32* d_prefixLen, u32* d_prefix, u32* d_child0, u32* d_child1, cudaStream_t s)
{
    if (!numNodes) return BM_OK;
    copy_topology_kernel<<<div_up(numNodes, 256), 256, 0, s>>>(g, d_childRefs, d_nodeMeta, numNodes, d_prefixLen, d_prefix, d_child0, d_child1);
    BM_LAUNCHED();
    return BM_OK;
}
int copy_child_boxes(const u32* d_childRefs, const float4* d_leafBox2, const float4* d_nodeBox2, u32 numNodes, float* d_boxes12, cudaStream_t s)
{
    if (!numNodes) return BM_OK;
    copy_child_boxes_kernel<<<div_up(numNodes, 256), 256, 0, s>>>(d_childRefs, d_leafBox2, d_nodeBox2, numNodes, d_boxes12);
    BM_LAUNCHED();
    return BM_OK;
}
int copy_wide_slots(const Grid& g, const uint4* d_wnodes, u32 numNodes, float* d_boxes24, u32* d_refs4, cudaStream_t s)
{
    if (!numNodes) return BM_OK;
    copy_wide_slots_kernel<<<div_up(numNodes, 256), 256, 0, s>>>(g, d_wnodes, numNodes, d_boxes24, d_refs4);
    BM_LAUNCHED();
    return BM_OK;
}
int canonical_keys(const u64* d_keys, const u32* d_vals, size_t n, int subBits, u64* d_out, cudaStream_t s)
{
    if (!n) return BM_OK;
    canonical_keys_kernel<<<div_up(n, 256), 256, 0, s>>>(d_keys, d_vals, n, subBits, d_out);
    BM_LAUNCHED();
    return BM_OK;
}
int split_canonical(const u64* d_keys, size_t n, u32* d_code, u32* d_tri, cudaStream_t s)
{
    if (!n) return BM_OK;
    split_canonical_kernel<<<div_up(n, 256), 256, 0, s>>>(d_keys, n, d_code, d_tri);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
