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
// sub-key of a pair: SUB_BITS-bit Morton code of the centre of (triangle box clipped to the cell) inside the cell
__device__ __forceinline__ u32 sub_key(const Grid& g, u32 ix, u32 iy, u32 iz, V3 tmin, V3 tmax)
{
    V3 mn, mx; cell_box(g, g.D, ix, iy, iz, mn, mx);
    float cx = .5f * (fmaxf(tmin.x, mn.x) + fminf(tmax.x, mx.x)), cy = .5f * (fmaxf(tmin.y, mn.y) + fminf(tmax.y, mx.y)),
          cz = .5f * (fmaxf(tmin.z, mn.z) + fminf(tmax.z, mx.z));
    constexpr int Q = 1 << (SUB_BITS / 3);
    int qx = (int)((cx - mn.x) / (mx.x - mn.x) * (float)Q), qy = (int)((cy - mn.y) / (mx.y - mn.y) * (float)Q), qz = (int)((cz - mn.z) / (mx.z - mn.z) * (float)Q);
    qx = min(max(qx, 0), Q - 1); qy = min(max(qy, 0), Q - 1); qz = min(max(qz, 0), Q - 1);
    return (spread3((u32)qx) << 2) | (spread3((u32)qy) << 1) | spread3((u32)qz);
}

struct WriteEmit
{
    u64* keys; u32* vals; u32 pos; u32 tri; const Grid* g; V3 tmin, tmax;
    __device__ void operator()(u32 code, u32 ix, u32 iy, u32 iz)
    {
        keys[pos] = ((u64)code << SUB_BITS) | sub_key(*g, ix, iy, iz, tmin, tmax); vals[pos] = tri; ++pos;
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
// check, a few need one or more SATs (~250 instructions each).  Done lane by lane the SATs of a warp execute one or two lanes at a
// time (first version: 4.5 of 32 lanes active, profiles/r1_final).  Here each lane only does the cheap margin checks; whenever a SAT
// is needed the (owner lane, candidate, level) item goes into a per-warp shared-memory ring, and the warp then drains the ring 32
// items at a time -- every lane runs one SAT, possibly for another lane's triangle (vertices are staged in shared memory), and
// continues that item's walk.  Same tests on the same boxes as pair_test, only scheduled differently.
constexpr int CL_WARPS = KG_THREADS / 32;
constexpr int CL_RING = 2048;                  // >= 32 lanes x 64 candidates: every candidate has at most one item alive

__device__ __forceinline__ void level_box(const Grid& g, int d, u32 ix, u32 iy, u32 iz, float bc[3], float hs[3])
{
    int sx = g.k[0] - axis_bits(d, 0), sy = g.k[1] - axis_bits(d, 1), sz = g.k[2] - axis_bits(d, 2);
    V3 mn, mx; cell_box(g, d, ix >> sx, iy >> sy, iz >> sz, mn, mx);
    bc[0] = (mx.x + mn.x) * .5f; bc[1] = (mx.y + mn.y) * .5f; bc[2] = (mx.z + mn.z) * .5f;       // BuildTree.cu:221-225
    hs[0] = (mx.x - mn.x) * .5f; hs[1] = (mx.y - mn.y) * .5f; hs[2] = (mx.z - mn.z) * .5f;
}
__device__ __forceinline__ bool vertex_inside(const TriSat& t, const float bc[3], const float hs[3], float m)
{
    float hx = hs[0] - m, hy = hs[1] - m, hz = hs[2] - m;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (fabsf(t.v[k].x - bc[0]) <= hx && fabsf(t.v[k].y - bc[1]) <= hy && fabsf(t.v[k].z - bc[2]) <= hz) return true;
    return false;
}

__global__ void __launch_bounds__(KG_THREADS) classify_kernel(Grid g, const float4* __restrict__ soup, u32 numTris, float margin,
                                                              u32* __restrict__ counts, u64* __restrict__ masks, u64* __restrict__ infos,
                                                              u32* __restrict__ largeList, u32* largeCount)
{
    __shared__ u32 ring[CL_WARPS][CL_RING];
    __shared__ float sv[CL_WARPS][32][9];
    __shared__ u32 sorg[CL_WARPS][32][4];               // x0, y0, z0, nx | ny << 8
    __shared__ unsigned long long smask[CL_WARPS][32];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t < numTris;
    float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
    if (valid) { const float4* p = soup + 3 * (size_t)t; a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); }
    TriSat ts; trisat_init(ts, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z));
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
    // stage this lane's triangle and candidate box for the other lanes
    sv[w][lane][0] = a.x; sv[w][lane][1] = a.y; sv[w][lane][2] = a.z; sv[w][lane][3] = b.x; sv[w][lane][4] = b.y; sv[w][lane][5] = b.z;
    sv[w][lane][6] = c.x; sv[w][lane][7] = c.y; sv[w][lane][8] = c.z;
    sorg[w][lane][0] = (u32)lo[0]; sorg[w][lane][1] = (u32)lo[1]; sorg[w][lane][2] = (u32)lo[2]; sorg[w][lane][3] = nx | (ny << 8);
    smask[w][lane] = 0ull;
    __syncwarp();

    // phase A: margin check of every candidate at leaf level; the rest is queued for a SAT at that level
    u32 head = 0, tail = 0;                            // warp-uniform ring cursors (monotone; index = cursor % CL_RING)
    u32 maxCand = __reduce_max_sync(FULL, ncand);
    for (u32 k = 0; k < maxCand; k++)
    {
        bool need = false;
        if (k < ncand)
        {
            u32 ix = (u32)lo[0] + k % nx, iy = (u32)lo[1] + (k / nx) % ny, iz = (u32)lo[2] + k / (nx * ny);
            float bc[3], hs[3]; level_box(g, g.D, ix, iy, iz, bc, hs);
            if (vertex_inside(ts, bc, hs, margin)) smask[w][lane] |= 1ull << k;
            else need = true;
        }
        unsigned bits = __ballot_sync(FULL, need);
        if (need) ring[w][(tail + __popc(bits & ((1u << lane) - 1u))) % CL_RING] = lane | (k << 5) | ((u32)g.D << 11);
        tail += __popc(bits);
    }
    __syncwarp();

    // phase B: drain the ring, 32 SATs per step
    while (head < tail)
    {
        u32 n = min(32u, tail - head);
        bool have = lane < n;
        u32 item = have ? ring[w][(head + lane) % CL_RING] : 0u;
        head += n;
        __syncwarp();
        bool again = false; u32 nextItem = 0;
        if (have)
        {
            u32 owner = item & 31u, k = (item >> 5) & 63u; int d = (int)(item >> 11);
            TriSat ot;
            trisat_init(ot, mk3(sv[w][owner][0], sv[w][owner][1], sv[w][owner][2]), mk3(sv[w][owner][3], sv[w][owner][4], sv[w][owner][5]),
                        mk3(sv[w][owner][6], sv[w][owner][7], sv[w][owner][8]));
            u32 onx = sorg[w][owner][3] & 255u, ony = sorg[w][owner][3] >> 8;
            u32 ix = sorg[w][owner][0] + k % onx, iy = sorg[w][owner][1] + (k / onx) % ony, iz = sorg[w][owner][2] + k / (onx * ony);
            float bc[3], hs[3]; level_box(g, d, ix, iy, iz, bc, hs);
            if (sat_box(ot, bc, hs))
            {
                // the walk goes on one level up: margin check there, else another SAT
                bool accept = d == 1;
                if (!accept)
                {
                    level_box(g, d - 1, ix, iy, iz, bc, hs);
                    if (vertex_inside(ot, bc, hs, margin)) accept = true;
                    else { again = true; nextItem = owner | (k << 5) | ((u32)(d - 1) << 11); }
                }
                if (accept) atomicOr(&smask[w][owner], 1ull << k);
            }
        }
        unsigned bits = __ballot_sync(FULL, again);
        if (again) ring[w][(tail + __popc(bits & ((1u << lane) - 1u))) % CL_RING] = nextItem;
        tail += __popc(bits);
        __syncwarp();
    }
    if (!valid) return;
    if (large)
    {
        // large triangle: handed to the warp-cooperative descent (large_kernel), which fills counts[t]
        counts[t] = 0; masks[t] = 0; infos[t] = 1ull << 63;
        largeList[atomicAdd(largeCount, 1u)] = t;
        return;
    }
    u64 mask = smask[w][lane];
    counts[t] = (u32)__popcll(mask);
    masks[t] = mask;
    infos[t] = empty ? 0ull : ((u64)lo[0] | ((u64)lo[1] << 11) | ((u64)lo[2] << 22) | ((u64)nx << 33) | ((u64)ny << 40));
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
        keys[pos] = ((u64)encode_path(g.D, ix, iy, iz) << SUB_BITS) | sub_key(g, ix, iy, iz, tmin, tmax);
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
                        keys[pos] = ((u64)encode_path(g.D, cx, cy, cz) << SUB_BITS) | sub_key(g, cx, cy, cz, tmin, tmax);
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
    classify_kernel<<<div_up(numTris, KG_THREADS), KG_THREADS, 0, s>>>(g, d_soup, numTris, margin, d_counts, d_masks, d_infos, d_largeList, d_largeCount);
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
