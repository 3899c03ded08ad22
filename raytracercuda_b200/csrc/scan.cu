// scan.cu -- device-wide exclusive prefix sum over u32 (pair offsets per triangle, leaf indices per pair, radix digit
// offsets per tile).  Reduce-then-scan over 4096-element tiles with warp-shuffle scans; tile sums are scanned recursively.
// HBM traffic: 2 reads + 1 write of the array (12 B/element).
#include "bm_device.cuh"

namespace bm
{
namespace
{
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 4096

__device__ __forceinline__ u32 warp_inclusive_scan(u32 v)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (unsigned)o) v += n; }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* total)
{
    __shared__ u32 warpSums[SCAN_THREADS / 32];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 inc = warp_inclusive_scan(v);
    if (lane == 31) warpSums[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        u32 s = lane < SCAN_THREADS / 32 ? warpSums[lane] : 0;
        u32 si = warp_inclusive_scan(s);
        if (lane < SCAN_THREADS / 32) warpSums[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) *total = si;
    }
    __syncthreads();
    u32 r = warpSums[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) tile_sums_kernel(const u32* __restrict__ in, size_t n, u32* __restrict__ sums)
{
    __shared__ u32 total;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    u32 s = 0;
    if (base + SCAN_ITEMS <= n && (((uintptr_t)(in + base)) & 15) == 0)
    {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
        uint4 a = p[0], b = p[1];
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += in[base + i];
    }
    (void)block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// scans one tile; adds tileOffsets[blockIdx.x] when given; optionally writes the grand total (single-tile case)
__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_kernel(const u32* in, u32* out, size_t n,   // in == out allowed
                                                                 const u32* __restrict__ tileOffsets, u32* __restrict__ totalOut)
{
    __shared__ u32 total;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    u32 ex = block_exclusive_scan(s, &total);
    if (tileOffsets) ex += tileOffsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    if (totalOut && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *totalOut = (tileOffsets ? tileOffsets[blockIdx.x] : 0) + total;
}
// small arrays (digit tables of small sorts, cluster counts ...): one block walks the array tile by tile carrying the running
// total -- one launch instead of three
__global__ void __launch_bounds__(SCAN_THREADS) small_scan_kernel(const u32* in, u32* out, size_t n, u32* __restrict__ totalOut)
{
    __shared__ u32 total;
    u32 carry = 0;
    for (size_t tileBase = 0; tileBase < n; tileBase += SCAN_TILE)
    {
        size_t base = tileBase + (size_t)threadIdx.x * SCAN_ITEMS;
        u32 v[SCAN_ITEMS], s = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
        u32 ex = block_exclusive_scan(s, &total) + carry;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
        carry += total;
        __syncthreads();
    }
    if (totalOut && threadIdx.x == 0) *totalOut = carry;
}
constexpr size_t SMALL_SCAN_MAX = 2 * (size_t)SCAN_TILE;     // beyond two tiles the serial walk loses to the three-kernel path (measured: 55 us vs 12 us at 63 K)
} // namespace

size_t scan_temp_bytes(size_t n)
{
    size_t bytes = 256, m = n;
    while (m > (size_t)SCAN_TILE) { m = (m + SCAN_TILE - 1) / SCAN_TILE; bytes += ((m * 4 + 255) / 256) * 256; }
    return bytes;
}

int exclusive_scan_u32(const u32* d_in, u32* d_out, size_t n, u32* d_total, void* d_temp, cudaStream_t s)
{
    if (n == 0) { if (d_total) BM_CUDA(cudaMemsetAsync(d_total, 0, 4, s)); return BM_OK; }
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n <= SMALL_SCAN_MAX)
    {
        small_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(d_in, d_out, n, d_total);
        BM_LAUNCHED();
        return BM_OK;
    }
    u32* sums = (u32*)d_temp;
    void* next = (char*)d_temp + ((tiles * 4 + 255) / 256) * 256;
    tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, n, sums);
    BM_LAUNCHED();
    BM_TRY(exclusive_scan_u32(sums, sums, tiles, d_total, next, s));
    // when d_total is requested the recursion wrote it (sum of tile sums == grand total)
    tile_scan_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, d_out, n, sums, nullptr);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
