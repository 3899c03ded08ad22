// sort.cu -- stable LSD radix sort of (leaf code, triangle id) pairs, 8 bits per pass, staged through shared memory with
// warp-level match/ballot ranking (north_star: "a device radix sort staged through shared memory, with warp-level scans
// and ballots").
//
// One-sweep form (default): ONE histogram pass over the input counts every digit of every pass, then one kernel per pass ranks a
// tile, obtains its per-digit global offsets by decoupled look-back over the preceding tiles, and scatters.  Three-kernel form (kept
// for >= 2^30 pairs and as BM_SORT=3k): tile histogram -> exclusive scan of the digit-major [256][tiles] table -> scatter.
// In both, a tile is 2048 (u64 keys) or 4096 (u32 keys) consecutive pairs handled by 8 warps; warp w owns a contiguous sub-range as
// rows of 32 lanes, so (warp, row, lane) order == input order and ranking inside a digit is stable; the tile is first permuted into
// digit order in shared memory, then written out so that consecutive threads write consecutive addresses of each digit run.
// HBM traffic per pass: 12 B read + 12 B written per pair (+ 8 B once for the histogram); 5 passes for 31-bit codes + 9-bit sub-keys.
#include <cstdlib>
#include <cstring>
#include "bm_device.cuh"

namespace bm
{
namespace
{
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;
// items per thread: 16 for 4-byte keys (tile 4096), 8 for 8-byte keys (tile 2048) -- keeps static shared memory under 48 KB
template <class K> struct RsCfg { static constexpr int ITEMS = sizeof(K) == 4 ? 16 : 8; static constexpr int TILE = RS_THREADS * ITEMS; };

template <class K>
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const K* __restrict__ keys, size_t n, int shift, u32 numTiles, u32* __restrict__ counts)
{
    constexpr int RS_ITEMS = RsCfg<K>::ITEMS, RS_TILE = RsCfg<K>::TILE;
    __shared__ u32 hist[RS_BINS];
    hist[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++)
    {
        size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
        u32 d = idx < n ? ((u32)(keys[idx] >> shift) & 0xFFu) : (u32)RS_BINS;
        // aggregate equal digits inside the warp before touching shared memory
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        if (d < RS_BINS && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[d], (u32)__popc(peers));
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * numTiles + blockIdx.x] = hist[threadIdx.x];
}

template <class K>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const K* __restrict__ keysIn, const u32* __restrict__ valsIn,
                                                                   K* __restrict__ keysOut, u32* __restrict__ valsOut,
                                                                   size_t n, int shift, u32 numTiles, const u32* __restrict__ scanned)
{
    constexpr int RS_ITEMS = RsCfg<K>::ITEMS, RS_TILE = RsCfg<K>::TILE;
    __shared__ u32 warpHist[RS_WARPS][RS_BINS + 1];     // +1: bin 256 swallows the out-of-range tail
    __shared__ u32 digitStart[RS_BINS + 1];
    __shared__ u32 globalBase[RS_BINS];
    __shared__ K sKeys[RS_TILE];
    __shared__ u32 sVals[RS_TILE];

    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * (RS_BINS + 1); i += RS_THREADS) (&warpHist[0][0])[i] = 0;
    __syncthreads();

    const size_t tileBase = (size_t)blockIdx.x * RS_TILE;
    const size_t warpBase = tileBase + (size_t)w * (RS_ITEMS * 32);
    K key[RS_ITEMS]; u32 val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
    {
        size_t idx = warpBase + (size_t)r * 32 + lane;
        bool ok = idx < n;
        key[r] = ok ? keysIn[idx] : (K)0;
        val[r] = ok ? valsIn[idx] : 0u;
    }
    // stable rank of every item among the items of the same digit in this warp
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
    {
        size_t idx = warpBase + (size_t)r * 32 + lane;
        u32 d = idx < n ? ((u32)(key[r] >> shift) & 0xFFu) : (u32)RS_BINS;
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        unsigned leader = __ffs(peers) - 1;
        u32 before = 0;
        if (lane == leader) { before = warpHist[w][d]; warpHist[w][d] = before + __popc(peers); }
        before = __shfl_sync(0xFFFFFFFFu, before, leader);
        rank[r] = before + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, tile total, global base of this tile's run
    {
        u32 d = threadIdx.x;      // RS_THREADS == RS_BINS
        u32 run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { u32 c = warpHist[ww][d]; warpHist[ww][d] = run; run += c; }
        digitStart[d] = run;      // tile count of digit d, scanned below
        globalBase[d] = scanned[(size_t)d * numTiles + blockIdx.x];
    }
    __syncthreads();
    if (w == 0)
    {
        // exclusive scan of the 256 digit counts by one warp (8 per lane)
        u32 c[8], s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { c[i] = digitStart[lane * 8 + i]; s += c[i]; }
        u32 inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (unsigned)o) inc += t; }
        u32 ex = inc - s;
#pragma unroll
        for (int i = 0; i < 8; i++) { digitStart[lane * 8 + i] = ex; ex += c[i]; }
        if (lane == 31) digitStart[RS_BINS] = ex;
    }
    __syncthreads();
    // permute the tile into digit order in shared memory
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
    {
        size_t idx = warpBase + (size_t)r * 32 + lane;
        if (idx < n)
        {
            u32 d = (u32)(key[r] >> shift) & 0xFFu;
            u32 pos = digitStart[d] + warpHist[w][d] + rank[r];
            sKeys[pos] = key[r];
            sVals[pos] = val[r];
        }
    }
    __syncthreads();
    const u32 valid = digitStart[RS_BINS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++)
    {
        u32 pos = i * RS_THREADS + threadIdx.x;
        if (pos < valid)
        {
            K k = sKeys[pos];
            u32 d = (u32)(k >> shift) & 0xFFu;
            size_t g = (size_t)globalBase[d] + (pos - digitStart[d]);
            keysOut[g] = k;
            valsOut[g] = sVals[pos];
        }
    }
}

// ---- one-sweep variant (Adinets & Merrill 2022): per pass ONE kernel.  The digit counts of every pass come from one upfront
// histogram of the input; a tile ranks its items, publishes its per-digit counts and obtains, per digit, the number of equal digits in
// the tiles before it by decoupled look-back (thread d walks digit d's chain).  Tiles are numbered by an atomic ticket in the order
// the blocks start, so a block only waits for blocks that already run; waits are bounded and trap.  status word: flag << 30 | count.
constexpr int RS_MAX_PASSES = 8;
// one-sweep tile: items per thread (the tile is staged in DYNAMIC shared memory, so it may exceed the 48 KB static limit)
#ifndef BM_OS_ITEMS64
#define BM_OS_ITEMS64 12
#endif
// (tiles of 1024 keys for small inputs were tried: a 503 K-pair sort takes 103 us with them, 94 us with the 3072-key tiles)
template <class K> struct OsCfg { static constexpr int ITEMS = sizeof(K) == 4 ? 16 : BM_OS_ITEMS64; static constexpr int TILE = RS_THREADS * ITEMS;
                                  static constexpr size_t SMEM = (size_t)TILE * (sizeof(K) + 4); };
constexpr u32 OS_AGG = 1u << 30, OS_INC = 2u << 30, OS_VAL = (1u << 30) - 1u;

template <class K>
__global__ void __launch_bounds__(RS_THREADS) radix_all_hist_kernel(const K* __restrict__ keys, size_t n, int passes, u32* __restrict__ hist)
{
    __shared__ u32 h[RS_MAX_PASSES][RS_BINS];
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_BINS; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    for (size_t idx = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; idx < n; idx += (size_t)gridDim.x * RS_THREADS)
    {
        K k = keys[idx];
        for (int p = 0; p < passes; p++) atomicAdd(&h[p][(u32)(k >> (8 * p)) & 0xFFu], 1u);
    }
    __syncthreads();
    for (int p = 0; p < passes; p++)
    {
        u32 c = h[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * RS_BINS + threadIdx.x], c);
    }
}

template <class K>
__global__ void __launch_bounds__(RS_THREADS) radix_onesweep_kernel(const K* __restrict__ keysIn, const u32* __restrict__ valsIn,
                                                                    K* __restrict__ keysOut, u32* __restrict__ valsOut, size_t n, int shift,
                                                                    const u32* __restrict__ passHist, u32* ticket, volatile u32* status, u32* ticketNext, u32* statusNext)
{
    constexpr int RS_ITEMS = OsCfg<K>::ITEMS, RS_TILE = OsCfg<K>::TILE;
    __shared__ u32 warpHist[RS_WARPS][RS_BINS + 1];     // +1: bin 256 swallows the out-of-range tail
    __shared__ u32 digitStart[RS_BINS + 1];
    __shared__ u32 globalBase[RS_BINS];
    __shared__ u32 scanTmp[RS_WARPS];
    extern __shared__ __align__(16) unsigned char osDyn[];
    K* sKeys = reinterpret_cast<K*>(osDyn);
    u32* sVals = reinterpret_cast<u32*>(osDyn + (size_t)RS_TILE * sizeof(K));
    __shared__ u32 sTile;

    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * (RS_BINS + 1); i += RS_THREADS) (&warpHist[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = sTile;
    // the NEXT pass's ticket and status words (the tables alternate; the other one was last used by the pass before this one, which has
    // ended): zeroed here, so that a pass is one kernel and nothing else on the stream
    statusNext[(size_t)tile * RS_BINS + threadIdx.x] = 0u;
    if (tile == 0 && threadIdx.x == 0) *ticketNext = 0u;

    const size_t tileBase = (size_t)tile * RS_TILE;
    const size_t warpBase = tileBase + (size_t)w * (RS_ITEMS * 32);
    K key[RS_ITEMS]; u32 val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
    {
        size_t idx = warpBase + (size_t)r * 32 + lane;
        bool ok = idx < n;
        key[r] = ok ? keysIn[idx] : (K)0;
        val[r] = ok ? valsIn[idx] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
    {
        size_t idx = warpBase + (size_t)r * 32 + lane;
        u32 d = idx < n ? ((u32)(key[r] >> shift) & 0xFFu) : (u32)RS_BINS;
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        unsigned leader = __ffs(peers) - 1;
        u32 before = 0;
        if (lane == leader) { before = warpHist[w][d]; warpHist[w][d] = before + __popc(peers); }
        before = __shfl_sync(0xFFFFFFFFu, before, leader);
        rank[r] = before + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {
        const u32 d = threadIdx.x;      // RS_THREADS == RS_BINS: thread d owns digit d
        u32 run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { u32 c = warpHist[ww][d]; warpHist[ww][d] = run; run += c; }
        digitStart[d] = run;            // tile count of digit d, scanned below
        // publish, then look back along digit d's chain
        volatile u32* st = status + d;
        u32 prefix = 0;
        if (tile == 0) st[0] = OS_INC | run;
        else
        {
            st[(size_t)tile * RS_BINS] = OS_AGG | run;
            for (long long t = (long long)tile - 1; t >= 0; --t)
            {
                u32 sv = st[(size_t)t * RS_BINS], spins = 0;
                while ((sv >> 30) == 0u) { if (++spins > (1u << 24)) __trap(); sv = st[(size_t)t * RS_BINS]; }
                prefix += sv & OS_VAL;
                if ((sv >> 30) == 2u) break;
            }
            st[(size_t)tile * RS_BINS] = OS_INC | (prefix + run);
        }
        // global base of digit d = number of keys with a smaller digit (exclusive scan of the pass histogram) + equal digits before this tile
        u32 hcount = passHist[d];
        u32 inc = hcount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (unsigned)o) inc += x; }
        if (lane == 31) scanTmp[w] = inc;
        __syncthreads();
        u32 woff = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) if ((unsigned)ww < w) woff += scanTmp[ww];
        globalBase[d] = woff + inc - hcount + prefix;
    }
    __syncthreads();
    if (w == 0)
    {
        u32 c[8], s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { c[i] = digitStart[lane * 8 + i]; s += c[i]; }
        u32 inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (unsigned)o) inc += t; }
        u32 ex = inc - s;
#pragma unroll
        for (int i = 0; i < 8; i++) { digitStart[lane * 8 + i] = ex; ex += c[i]; }
        if (lane == 31) digitStart[RS_BINS] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
    {
        size_t idx = warpBase + (size_t)r * 32 + lane;
        if (idx < n)
        {
            u32 d = (u32)(key[r] >> shift) & 0xFFu;
            u32 pos = digitStart[d] + warpHist[w][d] + rank[r];
            sKeys[pos] = key[r];
            sVals[pos] = val[r];
        }
    }
    __syncthreads();
    const u32 valid = digitStart[RS_BINS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++)
    {
        u32 pos = i * RS_THREADS + threadIdx.x;
        if (pos < valid)
        {
            K k = sKeys[pos];
            u32 d = (u32)(k >> shift) & 0xFFu;
            size_t g = (size_t)globalBase[d] + (pos - digitStart[d]);
            keysOut[g] = k;
            valsOut[g] = sVals[pos];
        }
    }
}
} // namespace

// smallest tile of any path and key width: sizes the per-tile tables
static constexpr size_t layout_tile()
{
    return (size_t)RS_THREADS * (BM_OS_ITEMS64 < 8 ? BM_OS_ITEMS64 : 8);
}

size_t sort_temp_bytes(size_t n)
{
    const size_t RS_TILE = layout_tile();         // the smallest tile any path uses -> the largest table
    size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    size_t counts = ((2 * tiles * RS_BINS * 4 + 255) / 256) * 256;  // per-tile digit table (three-kernel path) or two alternating tables of status words (one-sweep)
    return counts + scan_temp_bytes(tiles * RS_BINS) + RS_MAX_PASSES * RS_BINS * 4 + 256;
}

template <class K>
static int radix_sort_impl(K* d_keysA, u32* d_valsA, K* d_keysB, u32* d_valsB, size_t n, int bits, void* d_temp, cudaStream_t s, int* resultInB)
{
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    *resultInB = passes & 1;
    if (n == 0) return BM_OK;
    const size_t RS_TILE = RsCfg<K>::TILE;
    size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    const size_t layoutTiles = (n + layout_tile() - 1) / layout_tile();      // as sort_temp_bytes lays the buffer out
    u32* counts = (u32*)d_temp;
    void* scanTemp = (char*)d_temp + ((2 * layoutTiles * RS_BINS * 4 + 255) / 256) * 256;
    K *ka = d_keysA, *kb = d_keysB; u32 *va = d_valsA, *vb = d_valsB;
    static const bool threeKernel = getenv("BM_SORT") && strcmp(getenv("BM_SORT"), "3k") == 0;
    if (!threeKernel && n < (size_t)OS_VAL && passes <= RS_MAX_PASSES)
    {
        // layout: [status tiles*256 u32][scan temp (unused here)][hist passes*256 u32][ticket]
        u32* hist = (u32*)((char*)scanTemp + scan_temp_bytes(layoutTiles * RS_BINS));
        u32* ticket = hist + RS_MAX_PASSES * RS_BINS;                    // two tickets, two status tables: passes alternate between them
        BM_CUDA(cudaMemsetAsync(hist, 0, RS_MAX_PASSES * RS_BINS * 4 + 8, s));
        const size_t osTiles = (n + OsCfg<K>::TILE - 1) / OsCfg<K>::TILE;
        static std::atomic<int> smemSet[MAX_DEVICES];
        if (!smemSet[current_device()].load(std::memory_order_relaxed))
        {
            BM_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OsCfg<K>::SMEM));
            smemSet[current_device()].store(1, std::memory_order_relaxed);
        }
        unsigned hb = (unsigned)((n + RS_THREADS * 16 - 1) / (RS_THREADS * 16));
        unsigned cap = (unsigned)sm_count() * 8;
        radix_all_hist_kernel<K><<<hb < cap ? hb : cap, RS_THREADS, 0, s>>>(ka, n, passes, hist);
        BM_LAUNCHED();
        BM_CUDA(cudaMemsetAsync(counts, 0, osTiles * RS_BINS * 4, s));          // the first pass's table; every pass zeroes the next one's
        for (int p = 0; p < passes; p++)
        {
            u32* st = counts + (size_t)(p & 1) * osTiles * RS_BINS; u32* stNext = counts + (size_t)((p + 1) & 1) * osTiles * RS_BINS;
            radix_onesweep_kernel<K><<<(unsigned)osTiles, RS_THREADS, OsCfg<K>::SMEM, s>>>(ka, va, kb, vb, n, 8 * p, hist + p * RS_BINS, ticket + (p & 1), st, ticket + ((p + 1) & 1), stNext);
            BM_LAUNCHED();
            K* t = ka; ka = kb; kb = t;
            u32* tv = va; va = vb; vb = tv;
        }
        return BM_OK;
    }
    for (int p = 0; p < passes; p++)
    {
        int shift = 8 * p;
        radix_hist_kernel<K><<<(unsigned)tiles, RS_THREADS, 0, s>>>(ka, n, shift, (u32)tiles, counts);
        BM_LAUNCHED();
        BM_TRY(exclusive_scan_u32(counts, counts, tiles * RS_BINS, nullptr, scanTemp, s));
        radix_scatter_kernel<K><<<(unsigned)tiles, RS_THREADS, 0, s>>>(ka, va, kb, vb, n, shift, (u32)tiles, counts);
        BM_LAUNCHED();
        K* t = ka; ka = kb; kb = t;
        u32* tv = va; va = vb; vb = tv;
    }
    return BM_OK;
}

int radix_sort_pairs(u32* d_keysA, u32* d_valsA, u32* d_keysB, u32* d_valsB, size_t n, int bits, void* d_temp, cudaStream_t s, int* resultInB)
{
    return radix_sort_impl<u32>(d_keysA, d_valsA, d_keysB, d_valsB, n, bits, d_temp, s, resultInB);
}
int radix_sort_pairs64(u64* d_keysA, u32* d_valsA, u64* d_keysB, u32* d_valsB, size_t n, int bits, void* d_temp, cudaStream_t s, int* resultInB)
{
    return radix_sort_impl<u64>(d_keysA, d_valsA, d_keysB, d_valsB, n, bits, d_temp, s, resultInB);
}
} // namespace bm
