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
struct WriteEmit
{
    u64* keys; u32* vals; u32 pos; u32 tri; const Grid* g; V3 tmin, tmax;
    __device__ void operator()(u32 code, u32 ix, u32 iy, u32 iz)
    {
        V3 mn, mx; cell_box(*g, g->D, ix, iy, iz, mn, mx);
        float cx = .5f * (fmaxf(tmin.x, mn.x) + fminf(tmax.x, mx.x)), cy = .5f * (fmaxf(tmin.y, mn.y) + fminf(tmax.y, mx.y)),
              cz = .5f * (fmaxf(tmin.z, mn.z) + fminf(tmax.z, mx.z));
        int qx = (int)((cx - mn.x) / (mx.x - mn.x) * 32.f), qy = (int)((cy - mn.y) / (mx.y - mn.y) * 32.f), qz = (int)((cz - mn.z) / (mx.z - mn.z) * 32.f);
        qx = min(max(qx, 0), 31); qy = min(max(qy, 0), 31); qz = min(max(qz, 0), 31);
        u32 sub = (spread3((u32)qx) << 2) | (spread3((u32)qy) << 1) | spread3((u32)qz);
        keys[pos] = ((u64)code << SUB_BITS) | sub; vals[pos] = tri; ++pos;
    }
};

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
int emit_pairs(const Grid& g, const float4* d_soup, u32 numTris, const u32* d_offsets, u64* d_keys, u32* d_vals, cudaStream_t s)
{
    if (!numTris) return BM_OK;
    emit_kernel<<<div_up(numTris, KG_THREADS), KG_THREADS, 0, s>>>(g, d_soup, numTris, d_offsets, d_keys, d_vals);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
