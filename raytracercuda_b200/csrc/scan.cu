// scan.cu -- device-wide exclusive prefix sum over u32 (pair offsets per triangle, leaf indices per pair, radix digit
// offsets per tile).  One pass over 4096-element tiles (decoupled look-back between tiles, warp-shuffle scans inside);
// HBM traffic: 1 read + 1 write of the array (8 B/element), one launch.
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

// Single-pass scan with decoupled look-back (Merrill & Garland 2016): a tile publishes its aggregate, then sums the aggregates of
// its predecessors until it meets one that already knows its inclusive prefix.  Tiles are numbered by an atomic ticket in the
// order the blocks START, so a block only ever waits for blocks that are already running (no dependence on the block scheduler's
// order); the wait is bounded and traps instead of hanging.  status word = flag << 32 | value, flag 1 = aggregate, 2 = inclusive.
#define ST_AGG (1ull << 32)
#define ST_INC (2ull << 32)
// HITS: the input is an array of 16-byte hit records and the scanned value is "this record is a hit" (bounce-ray compaction without a
// separate flag pass)
template <bool HITS>
__global__ void __launch_bounds__(SCAN_THREADS) lookback_scan_kernel(const u32* in, u32* out, size_t n, u32* __restrict__ totalOut,   // in == out allowed
                                                                     u32* ticket, volatile unsigned long long* status, u32 numTiles, unsigned long long* total64)
{
    __shared__ u32 total, sTile, sPrefix;
    __shared__ unsigned long long wide;
    if (threadIdx.x == 0) { sTile = atomicAdd(ticket, 1u); wide = 0ull; }
    __syncthreads();
    const u32 tile = sTile;
    size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS], s = 0;
    if (HITS)
    {
        const uint4* h = reinterpret_cast<const uint4*>(in);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n && h[base + i].w != 0xFFFFFFFFu) ? 1u : 0u; s += v[i]; }
    }
    else if (base + SCAN_ITEMS <= n && (((uintptr_t)(in + base)) & 15) == 0)
    {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
        uint4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) s += v[i];
    }
    else
    {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    }
    u32 ex = block_exclusive_scan(s, &total);
    if (total64)
    {
        // the same sum in 64 bits, so that the caller can tell a total beyond 2^32 (the u32 prefixes have wrapped) from a valid one
        unsigned long long w = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) w += v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, o);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&wide, w);
        __syncthreads();
        if (threadIdx.x == 0 && wide) atomicAdd(total64, wide);
    }
    if (threadIdx.x < 32)
    {
        const unsigned FULL = 0xFFFFFFFFu, lane = threadIdx.x;
        const u32 agg = total;
        if (tile == 0)
        {
            if (lane == 0) { status[0] = ST_INC | agg; sPrefix = 0; }
        }
        else
        {
            if (lane == 0) status[tile] = ST_AGG | agg;
            u32 prefix = 0;
            long long look = (long long)tile - 1;
            for (;;)
            {
                long long idx = look - (long long)lane;
                unsigned long long st = idx >= 0 ? status[idx] : ST_INC;      // before the first tile: inclusive prefix 0
                u32 spins = 0;
                while (__any_sync(FULL, (st >> 32) == 0ull))
                {
                    if (++spins > (1u << 24)) __trap();                          // a predecessor never published: fail the launch, do not hang
                    st = idx >= 0 ? status[idx] : ST_INC;
                }
                unsigned incMask = __ballot_sync(FULL, (st >> 32) == 2ull);
                u32 val = (u32)st;
                if (incMask) { unsigned first = __ffs(incMask) - 1; if (lane > first) val = 0; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(FULL, val, o);
                prefix += val;
                if (incMask) break;
                look -= 32;
            }
            if (lane == 0) { status[tile] = ST_INC | (unsigned long long)(u32)(prefix + agg); sPrefix = prefix; }
        }
    }
    __syncthreads();
    ex += sPrefix;
    if (base + SCAN_ITEMS <= n && (((uintptr_t)(out + base)) & 15) == 0)
    {
        uint4 a, b;
        a.x = ex; ex += v[0]; a.y = ex; ex += v[1]; a.z = ex; ex += v[2]; a.w = ex; ex += v[3];
        b.x = ex; ex += v[4]; b.y = ex; ex += v[5]; b.z = ex; ex += v[6]; b.w = ex;
        uint4* q = reinterpret_cast<uint4*>(out + base);
        q[0] = a; q[1] = b;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    }
    if (totalOut && tile == numTiles - 1 && threadIdx.x == 0) *totalOut = sPrefix + total;
}

// small arrays (digit tables of small sorts, cluster counts ...): one block walks the array tile by tile carrying the running
// total -- one launch instead of three
__global__ void __launch_bounds__(SCAN_THREADS) small_scan_kernel(const u32* in, u32* out, size_t n, u32* __restrict__ totalOut, unsigned long long* total64)
{
    __shared__ u32 total;
    u32 carry = 0;
    unsigned long long wide = 0;
    for (size_t tileBase = 0; tileBase < n; tileBase += SCAN_TILE)
    {
        size_t base = tileBase + (size_t)threadIdx.x * SCAN_ITEMS;
        u32 v[SCAN_ITEMS], s = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
        u32 ex = block_exclusive_scan(s, &total) + carry;
        if (total64) { for (int i = 0; i < SCAN_ITEMS; i++) wide += v[i]; }
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
        carry += total;
        __syncthreads();
    }
    if (totalOut && threadIdx.x == 0) *totalOut = carry;
    if (total64 && wide) atomicAdd(total64, wide);
}
constexpr size_t SMALL_SCAN_MAX = 2 * (size_t)SCAN_TILE;     // beyond two tiles the serial walk loses (measured: 55 us at 63 K)
} // namespace

// temp: one status word per tile + the ticket
size_t scan_temp_bytes(size_t n)
{
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return tiles * 8 + 256;
}

int exclusive_scan_u32(const u32* d_in, u32* d_out, size_t n, u32* d_total, void* d_temp, cudaStream_t s, unsigned long long* d_total64)
{
    if (n == 0) { if (d_total) BM_CUDA(cudaMemsetAsync(d_total, 0, 4, s)); return BM_OK; }
    if (n <= SMALL_SCAN_MAX)
    {
        small_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(d_in, d_out, n, d_total, d_total64);
        BM_LAUNCHED();
        return BM_OK;
    }
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long* status = (unsigned long long*)d_temp;
    u32* ticket = (u32*)(status + tiles);
    BM_CUDA(cudaMemsetAsync(d_temp, 0, tiles * 8 + 8, s));
    lookback_scan_kernel<false><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, d_out, n, d_total, ticket, status, (u32)tiles, d_total64);
    BM_LAUNCHED();
    return BM_OK;
}

// out[i] = number of hits among hits[0..i-1]; *d_total = number of hits
int exclusive_scan_hit_flags(const bmHit* d_hits, u32* d_out, size_t n, u32* d_total, void* d_temp, cudaStream_t s)
{
    if (n == 0) { if (d_total) BM_CUDA(cudaMemsetAsync(d_total, 0, 4, s)); return BM_OK; }
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long* status = (unsigned long long*)d_temp;
    u32* ticket = (u32*)(status + tiles);
    BM_CUDA(cudaMemsetAsync(d_temp, 0, tiles * 8 + 8, s));
    lookback_scan_kernel<true><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(reinterpret_cast<const u32*>(d_hits), d_out, n, d_total, ticket, status, (u32)tiles, nullptr);
    BM_LAUNCHED();
    return BM_OK;
}
} // namespace bm
