// keygen.cu -- triangle soup gather and (leaf path code, triangle) pair generation.
//
// Replaces bmInsertTriangleInTree (BuildTree.cu:154-256): instead of inserting into a pointer-linked tree with atomicCAS /
// atomicAdd, every triangle emits the path codes of the leaf cells the reference would have put it in; the tree is then
// recovered from the sorted codes (hierarchy.cu).  Two passes (count, scan, emit) give every pair a deterministic slot, in
// triangle order, so a stable sort on the code alone yields the (code, triangle) order.
//
// Soup record: 3 x float4 per triangle = (v0.xyz, global id) (v1.xyz, mesh) (v2.xyz, face); read back with three 16-byte
// vector loads per thread (north_star: "vectorised float4 loads of the triangle soup").
#include "bm_device.cuh"
#include "bm_sat.cuh"

namespace bm
{
namespace
{
constexpr int KG_THREADS = 128;

__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ pos, const u32* __restrict__ idx, u32 numTris, u32 meshIdx, u32 base,
                                                     float4* __restrict__ soup)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numTris) return;
    u32 i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
    float4 a = make_float4(pos[3 * i0], pos[3 * i0 + 1], pos[3 * i0 + 2], __uint_as_float(base + t));
    float4 b = make_float4(pos[3 * i1], pos[3 * i1 + 1], pos[3 * i1 + 2], __uint_as_float(meshIdx));
    float4 c = make_float4(pos[3 * i2], pos[3 * i2 + 1], pos[3 * i2 + 2], __uint_as_float(t));
    float4* o = soup + 3 * (size_t)(base + t);
    o[0] = a; o[1] = b; o[2] = c;
}

struct NoEmit { BM_HD void operator()(u32, u32, u32, u32) const {} };

// Sort key of a pair = leaf path code (D bits) followed by a 15-bit Morton code of where the triangle sits INSIDE the leaf
// cell (centre of the triangle's box clipped to the cell, 5 bits per axis).  The code part decides everything parity is
// defined on; the sub-key only orders the triangles of one leaf spatially so that consecutive groups of four get tight
// cluster boxes (hierarchy.cu).  It is not part of any reference behaviour, so its arithmetic is free.
// sub-key of a pair: g.subBits-bit Morton code of the centre of (triangle box clipped to the cell) inside the cell
__device__ __forceinline__ u32 sub_key(const Grid& g, u32 ix, u32 iy, u32 iz, V3 tmin, V3 tmax)
{
    V3 mn, mx; cell_box(g, g.D, ix, iy, iz, mn, mx);
    float cx = .5f * (fmaxf(tmin.x, mn.x) + fminf(tmax.x, mx.x)), cy = .5f * (fmaxf(tmin.y, mn.y) + fminf(tmax.y, mx.y)),
          cz = .5f * (fmaxf(tmin.z, mn.z) + fminf(tmax.z, mx.z));
    const int Q = 1 << (g.subBits / 3);
    int qx = (int)((cx - mn.x) / (mx.x - mn.x) * (float)Q), qy = (int)((cy - mn.y) / (mx.y - mn.y) * (float)Q), qz = (int)((cz - mn.z) / (mx.z - mn.z) * (float)Q);
    qx = min(max(qx, 0), Q - 1); qy = min(max(qy, 0), Q - 1); qz = min(max(qz, 0), Q - 1);
    return (spread3((u32)qx) << 2) | (spread3((u32)qy) << 1) | spread3((u32)qz);
}

struct WriteEmit
{
    u64* keys; u32* vals; u32 pos; u32 tri; const Grid* g; V3 tmin, tmax;
    __device__ void operator()(u32 code, u32 ix, u32 iy, u32 iz)
    {
        keys[pos] = ((u64)code << g->subBits) | sub_key(*g, ix, iy, iz, tmin, tmax); vals[pos] = tri; ++pos;
    }
};

// ---- descent path (reference algorithm per triangle; used for triangles with more than 64 candidate cells, and as the
//      cross-check of the flat path when BM_KEYGEN=descent) ------------------------------------------------------------------
__global__ void __launch_bounds__(KG_THREADS) count_kernel(Grid g, const float4* __restrict__ soup, u32 numTris, u32* __restrict__ counts)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numTris) return;
    const float4* p = soup + 3 * (size_t)t;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    counts[t] = descend_triangle(g, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), NoEmit());
}

__global__ void __launch_bounds__(KG_THREADS) emit_kernel(Grid g, const float4* __restrict__ soup, u32 numTris, const u32* __restrict__ offsets,
                                                          u64* __restrict__ keys, u32* __restrict__ vals)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numTris) return;
    const float4* p = soup + 3 * (size_t)t;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    WriteEmit w = { keys, vals, offsets[t], t, &g,
                    mk3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z))),
                    mk3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z))) };
    descend_triangle<WriteEmit&>(g, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), w);
}

// ---- flat path: one SAT pass -------------------------------------------------------------------------------------------------
// classify: exact per-axis candidate ranges; up to 64 candidate cells -> decide each with pair_test's walk, remember the outcome as a
// bit mask so that the emit pass needs no geometry test at all; more than 64 -> queued for the warp-cooperative descent.
// info = ix0 | iy0 << 11 | iz0 << 22 | nx << 33 | ny << 40 | large << 63
//
// SIMT shape.  pair_test is a walk "margin check at level d, else SAT at level d, then d-1": most candidates end at the first margin
// check, a few need one or more SATs (~250 instructions each), and triangles have between 1 and 64 candidates.  Done lane by lane
// the SATs of a warp execute one or two lanes at a time and the candidate loops idle the lanes with few candidates (first version:
// 4.5 of 32 lanes active, profiles/r1_final).  Here a warp owns 32 * CL_ROUNDS triangles, staged in shared memory, and both kinds
// of work are spread over the lanes regardless of which triangle they belong to (measured on B200, 327 K-triangle mesh: one triangle
// per lane and four warps per block is fastest -- more triangles per warp fill the SAT steps better but leave too few warps):
//   * candidates: the (triangle slot, candidate) pairs of the warp are enumerated through a prefix sum of the candidate counts, 32
//     per step, each lane doing one leaf-level margin check;
//   * SATs: a candidate that needs one becomes an item (slot, candidate, level) in a per-warp ring; as soon as 32 items wait, every
//     lane runs one SAT and continues that item's walk (margin check one level up, else the item re-enters the ring).
// Same tests on the same boxes as pair_test, only scheduled differently.
#ifndef BM_CL_THREADS
#define BM_CL_THREADS 128
#endif
#ifndef BM_CL_ROUNDS
#define BM_CL_ROUNDS 1
#endif
#ifndef BM_CL_LONG
#define BM_CL_LONG 5
#endif
constexpr int CL_THREADS = BM_CL_THREADS;
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_ROUNDS = BM_CL_ROUNDS;
constexpr int CL_SLOTS = 32 * CL_ROUNDS;       // triangles per warp
constexpr int CL_LONG = BM_CL_LONG;                     // levels a walk runs item by item before the warp finishes it in one step
constexpr int CL_RING = 256;                   // never holds more than 31 + 32 + 32 items (drained whenever 32 wait)

__device__ __forceinline__ void level_box(const Grid& g, int d, u32 ix, u32 iy, u32 iz, float bc[3], float hs[3])
{
    int sx = g.k[0] - axis_bits(d, 0), sy = g.k[1] - axis_bits(d, 1), sz = g.k[2] - axis_bits(d, 2);
    V3 mn, mx; cell_box(g, d, ix >> sx, iy >> sy, iz >> sz, mn, mx);
    bc[0] = (mx.x + mn.x) * .5f; bc[1] = (mx.y + mn.y) * .5f; bc[2] = (mx.z + mn.z) * .5f;       // BuildTree.cu:221-225
    hs[0] = (mx.x - mn.x) * .5f; hs[1] = (mx.y - mn.y) * .5f; hs[2] = (mx.z - mn.z) * .5f;
}
// is one of the three vertices (9 floats) inside the box shrunk by m on every side?
__device__ __forceinline__ bool vertex_inside(const float* v9, const float bc[3], const float hs[3], float m)
{
    float hx = hs[0] - m, hy = hs[1] - m, hz = hs[2] - m;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (fabsf(v9[3 * k] - bc[0]) <= hx && fabsf(v9[3 * k + 1] - bc[1]) <= hy && fabsf(v9[3 * k + 2] - bc[2]) <= hz) return true;
    return false;
}

__global__ void __launch_bounds__(CL_THREADS) classify_kernel(Grid g, const float4* __restrict__ soup, u32 numTris, float margin,
                                                              u32* __restrict__ counts, u64* __restrict__ masks, u64* __restrict__ infos,
                                                              u32* __restrict__ largeList, u32* largeCount)
{
    __shared__ u32 ring[CL_WARPS][CL_RING];
    __shared__ float sv[CL_WARPS][CL_SLOTS][9];
    __shared__ u32 sorg[CL_WARPS][CL_SLOTS][4];         // x0, y0, z0, nx | ny << 8
    __shared__ u32 smask[CL_WARPS][CL_SLOTS][2];        // accepted candidates (bit k), as two words: 32-bit shared atomics are native
    __shared__ u32 spre[CL_WARPS][CL_SLOTS + 1];        // exclusive prefix of the candidate counts, slot order
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 tBase = (blockIdx.x * CL_WARPS + w) * CL_SLOTS;

    // ---- stage the warp's triangles: slot = round * 32 + lane ------------------------------------------------------------------------
    u32 carry = 0;
    unsigned largeBits = 0, emptyBits = 0;              // per-lane: bit r = triangle of round r is large / has no candidate
#pragma unroll
    for (int r = 0; r < CL_ROUNDS; r++)
    {
        u32 slot = r * 32 + lane, t = tBase + slot;
        bool valid = t < numTris;
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        if (valid) { const float4* p = soup + 3 * (size_t)t; a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); }
        int lo[3] = { 0, 0, 0 }, hi[3] = { -1, -1, -1 };
        if (valid)
        {
            axis_range(g, 0, fminf(a.x, fminf(b.x, c.x)), fmaxf(a.x, fmaxf(b.x, c.x)), lo[0], hi[0]);
            axis_range(g, 1, fminf(a.y, fminf(b.y, c.y)), fmaxf(a.y, fmaxf(b.y, c.y)), lo[1], hi[1]);
            axis_range(g, 2, fminf(a.z, fminf(b.z, c.z)), fmaxf(a.z, fmaxf(b.z, c.z)), lo[2], hi[2]);
        }
        bool empty = !valid || lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2];
        u32 nx = empty ? 0u : (u32)(hi[0] - lo[0] + 1), ny = empty ? 0u : (u32)(hi[1] - lo[1] + 1), nz = empty ? 0u : (u32)(hi[2] - lo[2] + 1);
        bool large = !empty && (u64)nx * ny * nz > 64ull;
        u32 ncand = (empty || large) ? 0u : nx * ny * nz;
        if (large) largeBits |= 1u << r;
        if (empty) emptyBits |= 1u << r;
        float* v = sv[w][slot];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = b.x; v[4] = b.y; v[5] = b.z; v[6] = c.x; v[7] = c.y; v[8] = c.z;
        sorg[w][slot][0] = (u32)lo[0]; sorg[w][slot][1] = (u32)lo[1]; sorg[w][slot][2] = (u32)lo[2]; sorg[w][slot][3] = nx | (ny << 8);
        smask[w][slot][0] = 0u; smask[w][slot][1] = 0u;
        u32 inc = ncand;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(FULL, inc, o); if (lane >= (unsigned)o) inc += x; }
        spre[w][slot] = carry + inc - ncand;
        carry += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) spre[w][CL_SLOTS] = carry;
    const u32 total = carry;
    __syncwarp();

    u32 head = 0, tail = 0;                            // warp-uniform ring cursors (monotone; index = cursor % CL_RING)
    // one SAT per lane on the next (up to) 32 ring items; an item whose walk goes on re-enters the ring.  A walk that is still
    // undecided CL_LONG levels above the leaf is almost always a triangle touching a major split plane (a vertex exactly on it: every
    // ancestor on the far side fails the margin check up to level 1) -- up to 30 more SATs in a row.  Those levels are independent
    // computations, so the warp evaluates them all at once, lane l taking level d - l, and replays the walk's decisions on the ballots.
    auto drain_step = [&]()
    {
        u32 n = min(32u, tail - head);
        bool have = lane < n;
        u32 item = have ? ring[w][(head + lane) % CL_RING] : 0u;
        head += n;
        __syncwarp();
        bool again = false; u32 nextItem = 0;
        if (have)
        {
            u32 slot = item & 127u, k = (item >> 7) & 63u; int d = (int)(item >> 13);
            const float* v = sv[w][slot];
            TriSat ot;
            trisat_init(ot, mk3(v[0], v[1], v[2]), mk3(v[3], v[4], v[5]), mk3(v[6], v[7], v[8]));
            u32 onx = sorg[w][slot][3] & 255u, ony = sorg[w][slot][3] >> 8;
            u32 ix = sorg[w][slot][0] + k % onx, iy = sorg[w][slot][1] + (k / onx) % ony, iz = sorg[w][slot][2] + k / (onx * ony);
            float bc[3], hs[3]; level_box(g, d, ix, iy, iz, bc, hs);
            if (sat_box(ot, bc, hs))
            {
                // the walk goes on one level up: margin check there, else another SAT
                bool accept = d == 1;
                if (!accept)
                {
                    level_box(g, d - 1, ix, iy, iz, bc, hs);
                    if (vertex_inside(v, bc, hs, margin)) accept = true;
                    else { again = true; nextItem = slot | (k << 7) | ((u32)(d - 1) << 13); }
                }
                if (accept) atomicOr(&smask[w][slot][k >> 5], 1u << (k & 31u));
            }
        }
        // long walks: all remaining levels at once
        unsigned longBits = __ballot_sync(FULL, again && (int)(nextItem >> 13) <= g.D - CL_LONG);
        if (again && (int)(nextItem >> 13) <= g.D - CL_LONG) again = false;
        while (longBits)
        {
            int src = __ffs(longBits) - 1;
            longBits &= longBits - 1;
            u32 it = __shfl_sync(FULL, nextItem, src);
            u32 slot = it & 127u, k = (it >> 7) & 63u; int d = (int)(it >> 13);       // state: SAT needed at level d (d <= 31: one lane per level)
            int e = d - (int)lane;
            bool valid = e >= 1, sOk = false, mOk = false;
            if (valid)
            {
                const float* v = sv[w][slot];
                TriSat ot;
                trisat_init(ot, mk3(v[0], v[1], v[2]), mk3(v[3], v[4], v[5]), mk3(v[6], v[7], v[8]));
                u32 onx = sorg[w][slot][3] & 255u, ony = sorg[w][slot][3] >> 8;
                u32 ix = sorg[w][slot][0] + k % onx, iy = sorg[w][slot][1] + (k / onx) % ony, iz = sorg[w][slot][2] + k / (onx * ony);
                float bc[3], hs[3]; level_box(g, e, ix, iy, iz, bc, hs);
                sOk = sat_box(ot, bc, hs);
                mOk = true;                                                              // level 1 passed its SAT: accepted
                if (e > 1) { level_box(g, e - 1, ix, iy, iz, bc, hs); mOk = vertex_inside(v, bc, hs, margin); }
            }
            unsigned failMask = __ballot_sync(FULL, valid && !sOk), accMask = __ballot_sync(FULL, valid && sOk && mOk);
            // the walk visits the lanes in order: the first lane that rejects (SAT fails) or accepts (SAT passes, then margin) decides
            bool accepted = accMask && (!failMask || __ffs(accMask) < __ffs(failMask));
            if (accepted && lane == 0) atomicOr(&smask[w][slot][k >> 5], 1u << (k & 31u));
        }
        unsigned bits = __ballot_sync(FULL, again);
        if (again) ring[w][(tail + __popc(bits & ((1u << lane) - 1u))) % CL_RING] = nextItem;
        tail += __popc(bits);
        __syncwarp();
    };

    // ---- candidates, 32 per step: leaf-level margin check; the rest is queued for a SAT at that level -----------------------------------
    for (u32 j0 = 0; j0 < total; j0 += 32)
    {
        u32 j = j0 + lane;
        bool need = false; u32 item = 0;
        if (j < total)
        {
            // slot of candidate j: the last slot whose prefix is <= j (slots without candidates share their successor's prefix)
            u32 loS = 0, hiS = CL_SLOTS;
#pragma unroll
            for (int it = 0; it < 7; it++) { u32 mid = (loS + hiS) >> 1; if (spre[w][mid] <= j) loS = mid; else hiS = mid; }
            while (spre[w][loS + 1] <= j) ++loS;       // (defensive; the search above already lands on the last such slot)
            u32 slot = loS, k = j - spre[w][slot];
            u32 nx = sorg[w][slot][3] & 255u, ny = sorg[w][slot][3] >> 8;
            u32 ix = sorg[w][slot][0] + k % nx, iy = sorg[w][slot][1] + (k / nx) % ny, iz = sorg[w][slot][2] + k / (nx * ny);
            float bc[3], hs[3]; level_box(g, g.D, ix, iy, iz, bc, hs);
            if (vertex_inside(sv[w][slot], bc, hs, margin)) atomicOr(&smask[w][slot][k >> 5], 1u << (k & 31u));
            else { need = true; item = slot | (k << 7) | ((u32)g.D << 13); }
        }
        unsigned bits = __ballot_sync(FULL, need);
        if (need) ring[w][(tail + __popc(bits & ((1u << lane) - 1u))) % CL_RING] = item;
        tail += __popc(bits);
        __syncwarp();
        while (tail - head >= 32u) drain_step();
    }
    while (head < tail) drain_step();
    __syncwarp();

    // ---- results ----------------------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < CL_ROUNDS; r++)
    {
        u32 slot = r * 32 + lane, t = tBase + slot;
        if (t >= numTris) continue;
        if (largeBits & (1u << r))
        {
            // large triangle: handed to the warp-cooperative descent (large_kernel), which fills counts[t]
            counts[t] = 0; masks[t] = 0; infos[t] = 1ull << 63;
            largeList[atomicAdd(largeCount, 1u)] = t;
            continue;
        }
        u64 mask = (u64)smask[w][slot][0] | ((u64)smask[w][slot][1] << 32);
        counts[t] = (u32)__popcll(mask);
        masks[t] = mask;
        u32 nxy = sorg[w][slot][3];
        infos[t] = (emptyBits & (1u << r)) ? 0ull : ((u64)sorg[w][slot][0] | ((u64)sorg[w][slot][1] << 11) | ((u64)sorg[w][slot][2] << 22) |
                                                     ((u64)(nxy & 255u) << 33) | ((u64)(nxy >> 8) << 40));
    }
}

__global__ void __launch_bounds__(KG_THREADS) emit_flat_kernel(Grid g, const float4* __restrict__ soup, u32 numTris, const u32* __restrict__ offsets,
                                                               const u64* __restrict__ masks, const u64* __restrict__ infos,
                                                               u64* __restrict__ keys, u32* __restrict__ vals)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numTris) return;
    u64 info = infos[t], mask = masks[t];
    if (!mask) return;
    const float4* p = soup + 3 * (size_t)t;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    V3 tmin = mk3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
    V3 tmax = mk3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
    u32 pos = offsets[t];
    if (info >> 63) return;                    // large triangle: emitted by large_kernel<true>
    u32 x0 = (u32)info & 0x7FFu, y0 = (u32)(info >> 11) & 0x7FFu, z0 = (u32)(info >> 22) & 0x7FFu, nx = (u32)(info >> 33) & 0x7Fu, ny = (u32)(info >> 40) & 0x7Fu;
    // the descent emits a triangle's codes in no particular order and the sort is keyed on (code, sub-key), so any order works here
    while (mask)
    {
        u32 bit = (u32)__ffsll((long long)mask) - 1u;
        mask &= mask - 1ull;
        u32 ix = x0 + bit % nx, iy = y0 + (bit / nx) % ny, iz = z0 + bit / (nx * ny);
        keys[pos] = ((u64)encode_path(g.D, ix, iy, iz) << g.subBits) | sub_key(g, ix, iy, iz, tmin, tmax);
        vals[pos] = t;
        ++pos;
    }
}
// ---- large triangles: warp-cooperative reference descent -----------------------------------------------------------------------------
// A triangle with more than 64 candidate cells (wings of f16, ground planes ...) would keep one thread busy for thousands of cells.
// Here a warp takes one such triangle: the frontier of kd nodes lives in a per-warp shared-memory stack, every iteration pops up to 32
// nodes, each lane evaluates the reference's test on the two children of its node (same expressions as descend_triangle: exact by
// construction), surviving children are pushed with a ballot-prefix, children at leaf depth are counted (pass 1) or written (pass 2).
constexpr int WD_WARPS = 4;
constexpr int WD_CAP = 1280;      // entries per warp; the LIFO frontier stays below 32 * (D + 1) + 64

template <bool EMIT>
__global__ void __launch_bounds__(WD_WARPS * 32) large_kernel(Grid g, const float4* __restrict__ soup, const u32* __restrict__ largeList,
                                                              const u32* __restrict__ largeCount, u32* __restrict__ counts, const u32* __restrict__ offsets,
                                                              u64* __restrict__ keys, u32* __restrict__ vals)
{
    __shared__ u64 stk[WD_WARPS][WD_CAP];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 nLarge = *largeCount;
    const u32 totalWarps = gridDim.x * WD_WARPS;
    u64* st = stk[w];
    for (u32 item = blockIdx.x * WD_WARPS + w; item < nLarge; item += totalWarps)
    {
        u32 t = largeList[item];
        const float4* p = soup + 3 * (size_t)t;
        float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        V3 va = mk3(a.x, a.y, a.z), vb = mk3(b.x, b.y, b.z), vc = mk3(c.x, c.y, c.z);
        V3 tmin = mk3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
        V3 tmax = mk3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
        int lo[3], hi[3];
        axis_range(g, 0, tmin.x, tmax.x, lo[0], hi[0]);
        axis_range(g, 1, tmin.y, tmax.y, lo[1], hi[1]);
        axis_range(g, 2, tmin.z, tmax.z, lo[2], hi[2]);
        TriSat ts; trisat_init(ts, va, vb, vc);
        __syncwarp();
        if (lane == 0) st[0] = 0;                 // root: depth 0, cell (0,0,0)
        __syncwarp();
        int top = 1;
        u32 total = 0, outPos = EMIT ? offsets[t] : 0u;
        while (top > 0)
        {
            int k = min(32, min(top, WD_CAP - top));
            u64 e = (int)lane < k ? st[top - 1 - lane] : 0ull;
            top -= k;
            __syncwarp();
            bool have = (int)lane < k;
            u32 d = (u32)(e >> 33), ix = (u32)(e >> 22) & 0x7FFu, iy = (u32)(e >> 11) & 0x7FFu, iz = (u32)e & 0x7FFu;
            int ax = d % 3;
            u32 nd = d + 1;
#pragma unroll
            for (int side = 0; side < 2; side++)
            {
                u32 cx = ix, cy = iy, cz = iz;
                if (ax == 0) cx = (ix << 1) | side; else if (ax == 1) cy = (iy << 1) | side; else cz = (iz << 1) | side;
                bool pass = have;
                if (pass)
                {
                    int shx = g.k[0] - axis_bits(nd, 0), shy = g.k[1] - axis_bits(nd, 1), shz = g.k[2] - axis_bits(nd, 2);
                    if ((int)(((cx + 1) << shx) - 1) < lo[0] || (int)(cx << shx) > hi[0]) pass = false;
                    else if ((int)(((cy + 1) << shy) - 1) < lo[1] || (int)(cy << shy) > hi[1]) pass = false;
                    else if ((int)(((cz + 1) << shz) - 1) < lo[2] || (int)(cz << shz) > hi[2]) pass = false;
                }
                if (pass)
                {
                    V3 mn, mx; cell_box(g, nd, cx, cy, cz, mn, mx);
                    float bc[3] = { (mx.x + mn.x) * .5f, (mx.y + mn.y) * .5f, (mx.z + mn.z) * .5f };      // BuildTree.cu:221-225
                    float hs[3] = { (mx.x - mn.x) * .5f, (mx.y - mn.y) * .5f, (mx.z - mn.z) * .5f };
                    pass = sat_box(ts, bc, hs) != 0;
                }
                bool leaf = (int)nd == g.D;              // warp-uniform only per lane's node depth: handle both kinds with two ballots
                unsigned leafBits = __ballot_sync(FULL, pass && leaf);
                unsigned pushBits = __ballot_sync(FULL, pass && !leaf);
                if (pass && leaf)
                {
                    if (EMIT)
                    {
                        u32 pos = outPos + __popc(leafBits & ((1u << lane) - 1u));
                        keys[pos] = ((u64)encode_path(g.D, cx, cy, cz) << g.subBits) | sub_key(g, cx, cy, cz, tmin, tmax);
                        vals[pos] = t;
                    }
                }
                outPos += __popc(leafBits); total += __popc(leafBits);
                if (pass && !leaf) st[top + __popc(pushBits & ((1u << lane) - 1u))] = ((u64)nd << 33) | ((u64)cx << 22) | ((u64)cy << 11) | (u64)cz;
                top += __popc(pushBits);
            }
            __syncwarp();
        }
        if (!EMIT && lane == 0) counts[t] = total;
    }
}
} // namespace

int gather_mesh(const float* d_pos, const u32* d_idx, u32 numTris, u32 meshIdx, u32 base, float4* d_soup, cudaStream_t s)
{
    if (!numTris) return BM_OK;
    gather_kernel<<<div_up(numTris, 256), 256, 0, s>>>(d_pos, d_idx, numTris, meshIdx, base, d_soup);
    BM_LAUNCHED();
    return BM_OK;
}
int count_pairs(const Grid& g, const float4* d_soup, u32 numTris, u32* d_counts, cudaStream_t s)
{
    if (!numTris) return BM_OK;
    count_kernel<<<div_up(numTris, KG_THREADS), KG_THREADS, 0, s>>>(g, d_soup, numTris, d_counts);
    BM_LAUNCHED();
    return BM_OK;
}
int classify_pairs(const Grid& g, const float4* d_soup, u32 numTris, float margin, u32* d_counts, u64* d_masks, u64* d_infos, u32* d_largeList,
                   u32* d_largeCount, cudaStream_t s)
{
    if (!numTris) return BM_OK;
    BM_CUDA(cudaMemsetAsync(d_largeCount, 0, 4, s));
    classify_kernel<<<div_up(numTris, CL_WARPS * CL_SLOTS), CL_THREADS, 0, s>>>(g, d_soup, numTris, margin, d_counts, d_masks, d_infos, d_largeList, d_largeCount);
    BM_LAUNCHED();
    // large triangles (their number stays on the device: the grid is fixed, warps loop over the list)
    large_kernel<false><<<(unsigned)sm_count() * 2, WD_WARPS * 32, 0, s>>>(g, d_soup, d_largeList, d_largeCount, d_counts, nullptr, nullptr, nullptr);
    BM_LAUNCHED();
    return BM_OK;
}
int emit_pairs_flat(const Grid& g, const float4* d_soup, u32 numTris, const u32* d_offsets, const u64* d_masks, const u64* d_infos, const u32* d_largeList,
                    const u32* d_largeCount, u64* d_keys, u32* d_vals, cudaStream_t s)
{
    if (!numTris) return BM_OK;
    emit_flat_kernel<<<div_up(numTris, KG_THREADS), KG_THREADS, 0, s>>>(g, d_soup, numTris, d_offsets, d_masks, d_infos, d_keys, d_vals);
    BM_LAUNCHED();
    large_kernel<true><<<(unsigned)sm_count() * 2, WD_WARPS * 32, 0, s>>>(g, d_soup, d_largeList, d_largeCount, nullptr, d_offsets, d_keys, d_vals);
    BM_LAUNCHED();
    return BM_OK;
}
int emit_pairs(const Grid& g, const float4* d_soup, u32 numTris, const u32* d_offsets, u64* d_keys, u32* d_vals, cudaStream_t s)
{
    if (!numTris) return BM_OK;
    emit_kernel<<<div_up(numTris, KG_THREADS), KG_THREADS, 0, s>>>(g, d_soup, numTris, d_offsets, d_keys, d_vals);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
