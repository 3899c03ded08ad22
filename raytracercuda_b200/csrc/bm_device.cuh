// bm_device.cuh -- device-side views, launch helpers and the internal interfaces between the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <string>
#include "bm_math.cuh"
#include "../../include/beam_b200.h"

namespace bm
{
// ---- error plumbing -------------------------------------------------------------------------------------------------
void set_error(const std::string& s);
const char* last_error();
#define BM_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { ::bm::set_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); return BM_ERROR_CUDA; } } while (0)
// after every kernel launch: count it (bmLaunchCount, the bench's gpu_launches evidence) and surface launch errors
extern std::atomic<unsigned long long> g_launches;
#define BM_LAUNCHED() do { ::bm::g_launches.fetch_add(1, std::memory_order_relaxed); BM_CUDA(cudaGetLastError()); } while (0)
#define BM_TRY(expr) do { int r_ = (expr); if (r_ != BM_OK) return r_; } while (0)

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
constexpr int MAX_DEVICES = 64;
int current_device();        // cudaGetDevice, clamped to [0, MAX_DEVICES)
int sm_count();              // of the current device (cached per device)
void set_sticky(int code, const char* where);

// ---- HBM layout of a finished tree (all index-based, no absolute pointers inside the arrays) --------------------------
//
//  nodes    : N = L-1 internal nodes of the binary radix tree over the L distinct leaf codes, 128 bytes each (two 64-byte
//             lines, 128-byte aligned):
//               float4 q0 = (lo0.x lo0.y lo0.z hi0.x)     child 0 = lower half ("left"); tight box = union of whole triangles
//               float4 q1 = (hi0.y hi0.z lo1.x lo1.y)
//               float4 q2 = (lo1.z hi1.x hi1.y hi1.z)
//               uint4  q3 = (child0, child1, prefix, prefixLen)   child bit 31 = leaf; prefix = top prefixLen bits of the codes
//               float4 q4 = (cellMin.xyz, split plane s)           the reference kd node at depth prefixLen IS this node: its
//               float4 q5 = (cellMax.xyz, split axis)              implicit box, axis prefixLen % 3 and plane .5f*(min+max)
//               q6, q7    = unused
//  leaves   : 48 bytes each, sorted by code: uint4 (code, first pair, pair count, first cluster), float4 (cellMin.xyz, cellMax.x),
//             float4 (cellMax.y, cellMax.z, -, -).
//  ltris    : one 48-byte record per (leaf, triangle) pair; the pairs of a leaf are contiguous and ordered by the in-cell
//             Morton sub-key:  float4 (v0.xyz, global triangle id) (e1.xyz = v1-v0, mesh index | EXCLUDED bit) (e2.xyz = v2-v0,
//             face index in mesh).  EXCLUDED marks pairs beyond the reference's 256-faces-per-leaf cap (BuildTree.cuh:17).
//  clusters : 32 bytes per group of CLUSTER consecutive pairs of one leaf: float4 (lo.xyz, hi.x) (hi.y, hi.z, -, -) = padded
//             union of the group's whole triangles.  A leaf tests a group's triangles only if the ray's line meets its box.
// in-cell sub-key width (Grid::subBits): 9 bits keep the sort key at 40 bits = 5 radix passes; meshes much finer than the leaf cells
// (hundreds of faces per leaf) get 15 bits, which tightens the 4-face cluster boxes inside such leaves (tree.cu picks per build)
constexpr int SUB_BITS_COARSE = 9, SUB_BITS_FINE = 15;
#ifndef BM_CLUSTER
#define BM_CLUSTER 4
#endif
constexpr int CLUSTER = BM_CLUSTER;          // faces per cluster box (a power of two <= 32)
constexpr int NODE_F4 = 8, LEAF_F4 = 3, CLUSTER_F4 = 2;
static const u32 LEAF_BIT = 0x80000000u;
static const u32 EXCLUDED_BIT = 0x80000000u;
struct TreeView
{
    const float4* nodes;      // 128-byte binary records (null unless the tree was built with BM_NODE_LAYOUT=both|binary)
    const uint4* wnodes;      // 64-byte wide records (bm_wide.cuh)
    const float4* leaves;
    const float4* ltris;
    const float4* clusters;
    u32 numLeaves;
    const u32* root;      // device: index of the root node (numLeaves >= 2; a single leaf is its own root)
    Grid grid;            // the wide records store cell coordinates: boxes and planes are rebuilt from the grid
    float leafExt[3];     // largest leaf-cell extent per axis (+ the boxes' pad): reach of a leaf cell beyond the triangles it holds
    float cellPad;        // CLOSEST mode culls with cells grown by this much (world extent * 2^-18: far above the rounding of a hit point)
};

// ---- scan.cu --------------------------------------------------------------------------------------------------------
size_t scan_temp_bytes(size_t n);
// out[i] = sum(in[0..i-1]); *d_total (device, nullable) = sum of all. in == out allowed. temp: scan_temp_bytes(n) bytes.
// d_total64 (device, nullable, must be zero before the call) receives the same total accumulated in 64 bits: a value >= 2^32 means the
// u32 prefixes have wrapped
int exclusive_scan_u32(const u32* d_in, u32* d_out, size_t n, u32* d_total, void* d_temp, cudaStream_t s, unsigned long long* d_total64 = nullptr);
// out[i] = number of hit records (tri != 0xFFFFFFFF) among d_hits[0..i-1]; *d_total = their number. Same temp as above.
int exclusive_scan_hit_flags(const bmHit* d_hits, u32* d_out, size_t n, u32* d_total, void* d_temp, cudaStream_t s);

// ---- sort.cu --------------------------------------------------------------------------------------------------------
size_t sort_temp_bytes(size_t n);
// Stable LSD radix sort of (key, value) pairs on the low `bits` bits of the key. Input in buffer A; A and B ping-pong;
// *resultInB tells which pair of buffers holds the sorted result (B after an odd number of 8-bit passes).
int radix_sort_pairs(u32* d_keysA, u32* d_valsA, u32* d_keysB, u32* d_valsB, size_t n, int bits, void* d_temp, cudaStream_t s, int* resultInB);
int radix_sort_pairs64(u64* d_keysA, u32* d_valsA, u64* d_keysB, u32* d_valsB, size_t n, int bits, void* d_temp, cudaStream_t s, int* resultInB);

// ---- keygen.cu ------------------------------------------------------------------------------------------------------
int gather_mesh(const float* d_pos, const u32* d_idx, u32 numTris, u32 meshIdx, u32 base, float4* d_soup, cudaStream_t s);
int count_pairs(const Grid& g, const float4* d_soup, u32 numTris, u32* d_counts, cudaStream_t s);
int emit_pairs(const Grid& g, const float4* d_soup, u32 numTris, const u32* d_offsets, u64* d_keys, u32* d_vals, cudaStream_t s);
// flat path: classify = candidate ranges + per-candidate test into a bit mask (one SAT pass), emit = decode the masks
int classify_pairs(const Grid& g, const float4* d_soup, u32 numTris, float margin, u32* d_counts, u64* d_masks, u64* d_infos, u32* d_largeList,
                   u32* d_largeCount, cudaStream_t s);
int emit_pairs_flat(const Grid& g, const float4* d_soup, u32 numTris, const u32* d_offsets, const u64* d_masks, const u64* d_infos, const u32* d_largeList,
                    const u32* d_largeCount, u64* d_keys, u32* d_vals, cudaStream_t s);

// ---- hierarchy.cu ---------------------------------------------------------------------------------------------------
int mark_heads(const u64* d_keys, size_t n, int subBits, u32* d_flags, cudaStream_t s);
int write_leaves(const Grid& g, const u64* d_keys, const u32* d_flags, const u32* d_leafIdx, size_t n, u32 numLeaves, float4* d_leaves, u32* d_clusterCounts,
                 u32* d_codes, cudaStream_t s);
int pack_leaf_tris(const float4* d_soup, const u32* d_vals, size_t n, float4* d_ltris, cudaStream_t s);
// cluster boxes + leaf boxes + leaf statistics + the 256-cap exclusion flags; d_clusterFirst = exclusive scan of the cluster counts
int leaf_bounds(float4* d_leaves, u32 numLeaves, const u32* d_flags, const u32* d_leafIdx, size_t numPairs, const u32* d_clusterFirst, float4* d_ltris,
                float4* d_clusters, float4* d_leafBox2, float pad, u32* d_stats, u32* d_overList, cudaStream_t s);
// radix tree + tight boxes + reference-node data in one bottom-up pass; *d_root receives the root's node index
// d_meet: scratch, 8 bytes per node (where the two children of a node meet); d_nodes (nullable): the 128-byte binary records; d_wnodes: the 64-byte wide records (bm_wide.cuh); d_nodeMeta: scratch, 8 bytes per node
int build_hierarchy(const Grid& g, const u32* d_codes, u32 numLeaves, const float4* d_leafBox2, float4* d_nodeBox2, unsigned long long* d_meet, u32* d_childRefs,
                    float4* d_nodes, uint4* d_wnodes, uint2* d_nodeMeta, u32* d_root, cudaStream_t s);
int copy_topology(const Grid& g, const u32* d_childRefs, const uint2* d_nodeMeta, u32 numNodes, u32* d_prefixLen, u32* d_prefix, u32* d_child0, u32* d_child1, cudaStream_t s);
int copy_child_boxes(const u32* d_childRefs, const float4* d_leafBox2, const float4* d_nodeBox2, u32 numNodes, float* d_boxes12, cudaStream_t s);
int copy_wide_slots(const Grid& g, const uint4* d_wnodes, u32 numNodes, float* d_boxes24, u32* d_refs4, cudaStream_t s);
int canonical_keys(const u64* d_keys, const u32* d_vals, size_t n, int subBits, u64* d_out, cudaStream_t s);
int split_canonical(const u64* d_keys, size_t n, u32* d_code, u32* d_tri, cudaStream_t s);

// ---- trace.cu -------------------------------------------------------------------------------------------------------
// sample == 0xFFFFFFFF: the exact pixel-centre table entry (what bmMarch traces); otherwise the direction is jittered inside the pixel
// (jitterX/Y = pixel size in units of zoom) by a counter-based hash of (ray, sample, seed) -- see camera_ray below
struct CameraIn { const float* initialRays; float eye[3]; float orient[9]; float jitterX, jitterY; u32 sample, seed; };
inline CameraIn make_camera(const float* d_initialRays3, const float* eye3, const float* orient9, float jx = 0.f, float jy = 0.f, u32 sample = 0xFFFFFFFFu, u32 seed = 0)
{
    CameraIn c; c.initialRays = d_initialRays3;
    for (int i = 0; i < 3; i++) c.eye[i] = eye3 ? eye3[i] : 0.f;
    for (int i = 0; i < 9; i++) c.orient[i] = orient9 ? orient9[i] : 0.f;
    c.jitterX = jx; c.jitterY = jy; c.sample = sample; c.seed = seed;
    return c;
}
#if defined(__CUDACC__)
__device__ __forceinline__ u32 pcg_hash(u32 v)
{
    u32 state = v * 747796405u + 2891336453u;
    u32 word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
__device__ __forceinline__ float u01(u32 h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
// Ray i of a camera: dir = orient * initialRays[i] from eye (glm mat3 * vec3, type_mat3x3.inl:424-430; Camera.cpp:85-97), optionally
// jittered.  ONE definition used by the trace kernels (which form camera rays in registers), bmCameraRays (which materialises them) and
// the bounce generator (which re-forms the primary ray of a hit), so all three see the same bits.
__device__ __forceinline__ void camera_ray(const CameraIn& cam, u64 i, u32 sample, V3& o, V3& d)
{
    const float* r = cam.initialRays + 3 * i;
    float x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
    if (sample != 0xFFFFFFFFu)
    {
        // jitter inside the pixel: the table stores unit vectors (rx,ry,zoom)*d; perturb in the image plane and renormalise
        u32 h = pcg_hash((u32)i ^ pcg_hash(sample * 0x9E3779B9u + cam.seed));
        float jx = (u01(h) - .5f) * cam.jitterX, jy = (u01(pcg_hash(h)) - .5f) * cam.jitterY;
        float s = 1.f / z;                       // z = zoom*d > 0
        float px = x * s + jx, py = y * s + jy;  // in units of zoom
        float inv = rsqrtf(1.f + px * px + py * py);
        x = px * inv; y = py * inv; z = inv;
    }
    const float* m = cam.orient;
    d = mk3(m[0] * x + m[3] * y + m[6] * z, m[1] * x + m[4] * y + m[7] * z, m[2] * x + m[5] * y + m[8] * z);
    o = mk3(cam.eye[0], cam.eye[1], cam.eye[2]);
}
__device__ __forceinline__ void camera_ray(const CameraIn& cam, u64 i, V3& o, V3& d) { camera_ray(cam, i, cam.sample, o, d); }

// The diffuse bounce ray of a hit (rays.cu): origin = hit point + 1e-4 n, direction cosine-weighted about the geometric normal n (flipped
// towards the side the incoming ray came from), random numbers = counter-based hash of (ray index, sample, bounce, seed).  ONE definition
// used by bounce_kernel (which writes the ray into a list) and by the trace kernels' bounce launches (which form it in registers).
__device__ __forceinline__ void bounce_ray(V3 ro, V3 rd, float t, u32 tri, const float4* __restrict__ soup, u64 i, u32 sample, u32 bounce, u32 seed, V3& o, V3& d)
{
    const float4* p = soup + 3 * (size_t)tri;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    V3 e1 = mk3(b.x - a.x, b.y - a.y, b.z - a.z), e2 = mk3(c.x - a.x, c.y - a.y, c.z - a.z);
    V3 nrm = cross3(e1, e2);
    float len2 = dot3(nrm, nrm);
    nrm = len2 > 0.f ? nrm * rsqrtf(len2) : mk3(0.f, 0.f, 1.f);
    if (dot3(nrm, rd) > 0.f) nrm = nrm * -1.f;                 // face the side the ray came from
    V3 hp = mk3(ro.x + t * rd.x, ro.y + t * rd.y, ro.z + t * rd.z);
    u32 h = pcg_hash((u32)i ^ pcg_hash(seed ^ pcg_hash(sample * 0x9E3779B9u + bounce * 0x85EBCA6Bu)));
    float u1 = u01(h), u2 = u01(pcg_hash(h ^ 0xB5297A4Du));
    float rr = sqrtf(u1), phi = 6.28318530718f * u2;
    float lx = rr * cosf(phi), ly = rr * sinf(phi), lz = sqrtf(fmaxf(0.f, 1.f - u1));
    // branchless orthonormal basis (Duff et al. 2017)
    float sgn = copysignf(1.f, nrm.z);
    float aa = -1.f / (sgn + nrm.z), bb = nrm.x * nrm.y * aa;
    V3 t1 = mk3(1.f + sgn * nrm.x * nrm.x * aa, sgn * bb, -sgn * nrm.x), t2 = mk3(bb, sgn + nrm.y * nrm.y * aa, -nrm.y);
    d = mk3(t1.x * lx + t2.x * ly + nrm.x * lz, t1.y * lx + t2.y * ly + nrm.y * lz, t1.z * lx + t2.z * ly + nrm.z * lz);
    o = mk3(hp.x + 1e-4f * nrm.x, hp.y + 1e-4f * nrm.y, hp.z + 1e-4f * nrm.z);
}
#endif
// Bounce launches (bmTraceBounceCamera): the camera launch's slots are the pixels of `samples` consecutive samples of a frame; a lane whose
// pixel has a primary hit (hits1[sample * n + pixel]) forms that hit's diffuse bounce ray in registers and traces it, the others stay idle.
// No compaction, no ray list: the rays of a warp are the hits of one 8x4 tile, and the launch runs under its own temporal tile schedule.
// accum / covered (nullable): the progressive estimator is applied where the ray retires (accum[pixel] += ambient + (escaped ? weight : 0)).
struct BounceIn { const bmHit* hits1; const float4* soup; u32 sample, bounce, seed;      // sample = index of the launch's first sample in the bounce rays' random sequence
                  float* accum; u32* covered; float ambient, weight; unsigned long long* rayCount; };
// d_workCounter: one u32 of device scratch per launch in flight (the persistent warps' work queue head)
// d_numRays (nullable): the ray count is read from device memory at kernel start (<= n, which then only sizes the grid)
int trace_rays(const TreeView& tv, const bmRay* d_rays, u64 n, const u32* d_numRays, int mode, bmHit* d_hits, u64* d_counters, u32* d_workCounter, cudaStream_t s);
// Temporal tile scheduling of camera launches (nullable): the kernel records what every 8x4 pixel tile cost, and the next launch of
// the same frame size starts the expensive tiles first (longest-processing-time-first), so that the handful of long silhouette rays
// overlaps with the bulk of the frame instead of forming its tail.  Any order is a permutation of the same tiles: results do not depend on it.
//   cost[tiles]  cost of each tile in the last launch          order[4*tiles]  work list of the next launch (valid when state[TS_OK] != 0):
//                tiles by descending cost, the most expensive ones as four quarter entries; state[TS_ENTRIES] = its length
//   state: [0, 64) histogram of cost buckets, [64, 128) scatter cursors, [TS_SUM] 64-bit total cost, [TS_WORK] the launch's work counter
//          (all zeroed by ONE memset per launch), [TS_OK] and [TS_ENTRIES] set by the ordering kernel.  cost[] needs no memset: the
//          ordering kernel, its only reader, zeroes every entry after reading it (zeroed once when the buffers are set up)
constexpr int TS_BUCKETS = 64, TS_SUM = 128, TS_WORK = 130, TS_OK = 132, TS_ENTRIES = 133, TS_WORDS = 134;       // [0, TS_WORK] is zeroed before every launch
//   hitFlag (nullable, camera launches): hitFlag[tile] = 1 where some ray of the tile hit
//   filter  (nullable, bounce launches): the hitFlag array of the camera launch whose hits this launch follows up -- the work list is then
//                made just BEFORE the launch: tiles that pass the filter, by descending cost in the previous bounce launch
struct TileSched { u32* cost; u32* order; u32* state; u32 tiles; u32* hitFlag; const u32* filter; };
// samples > 1: the launch covers samples cam.sample .. cam.sample + samples - 1 of every pixel (jitter applied in the kernel); d_hits then holds
// samples * n records, sample-major.  bounce (nullable): a bounce launch (BounceIn).
int trace_camera(const TreeView& tv, const CameraIn& cam, u64 n, u32 width, int mode, bmHit* d_hits, u32* d_pixels, const bmStaticMeshData* d_meshData,
                 u64* d_counters, u32* d_workCounter, const TileSched* sched, cudaStream_t s, u32 samples = 1, const BounceIn* bounce = nullptr);
bool tile_sched_usable();
int trace_rays_tiled(const TreeView& tv, const bmRay* d_rays, u64 n, u32 width, int mode, bmHit* d_hits, u64* d_counters, u32* d_workCounter,
                     const TileSched* sched, cudaStream_t s);
// Segments of a ray batch traced brick by brick (bricks.cu): keys[i] = position of the brick in the ray's visiting order << 24 | wave << bits | brick,
// rays[i] = ray index, sorted by (wave, brick); roots[brick] = reference of the brick's subtree; rayKey[ray] = the ray's merged result.
struct SegIn { const u32* keys; const u32* rays; const u32* roots; unsigned long long* rayKey; u32 brickMask; u32 brickShift; };      // brick = (key >> brickShift) & brickMask
int trace_segments(const TreeView& tv, const bmRay* d_rays, u64 numSegs, int mode, const SegIn& seg, u64* d_counters, u32* d_workCounter, cudaStream_t s);
int clear_u32(u32* d_buf, size_t n, u32 value, cudaStream_t s);
int shade_hits(const bmHit* d_hits, u64 n, const float4* d_soup, const bmStaticMeshData* d_meshData, const bmMaterial* d_materials, const bmTexture* d_textures,
               u32 numTextures, u32* d_pixels, cudaStream_t s);

// ---- bricks.cu ------------------------------------------------------------------------------------------------------
// Ray batches against trees far larger than the L2 (bmTraceRays): the top `brickBits` levels of the reference's kd-tree are an implicit
// grid of bricks; every ray is cut into one segment per brick it enters, the segments are sorted by brick and traced brick by brick, so
// that the subtree a warp walks is cache-resident while it is walked.  Results are identical to the direct traversal.
constexpr u32 BRICK_EMPTY = 0xFFFFFFFFu;
constexpr int BRICK_MAX_BITS = 15;
// roots[2^bits]: reference of the subtree holding exactly the leaves whose code starts with the brick's bits (BRICK_EMPTY: none);
// occ[2^(bits+1)]: heap over the levels above (occ[(1 << d) + path] != 0: some leaf below the depth-d node `path`)
int brick_table(const TreeView& tv, int bits, u32* d_roots, unsigned char* d_occ, cudaStream_t s);
int brick_coherence_bits(int bits);      // low key bits that order the segments of one brick by entry point and direction octant (sort on bits + 2 + these)
// one entry per (ray, brick entered), in the ray's visiting order: keys = order << 24 | wave << bits | brick, rays = ray index.  *d_count receives the
// number of entries the batch needs; entries beyond `capacity` are not written.
int brick_segments(const TreeView& tv, int bits, const u32* d_roots, const unsigned char* d_occ, const bmRay* d_rays, u64 n, int mode,
                   u32* d_keys, u32* d_rayIdx, u32 capacity, unsigned long long* d_count, cudaStream_t s);
// rayKey -> hit records: t, u, v recomputed from the winning triangle with the leaf records' arithmetic
int brick_resolve(const bmRay* d_rays, u64 n, int mode, const unsigned long long* d_rayKey, const float4* d_soup, bmHit* d_hits, cudaStream_t s);

// ---- rays.cu --------------------------------------------------------------------------------------------------------
int camera_rays(const CameraIn& cam, u64 n, bmRay* d_out, cudaStream_t s);
int add_count(const u32* d_count, u64 add, u64* d_total, cudaStream_t s);      // *d_total += *d_count + add (ray accounting without a host round trip)
int bounce_rays(const bmRay* d_rays, const CameraIn* cam, const bmHit* d_hits, u64 n, const float4* d_soup, u32 sample, u32 bounce, u32 seed,
                u32* d_flagsAndOffsets, void* d_scanTemp, bmRay* d_out, u32* d_srcIndex, u32* d_count, cudaStream_t s, u32* d_appendTotal = nullptr);
int progressive_accumulate_bounce(const bmHit* d_hits2, const u32* d_src2, const u32* d_count2, u64 maxRays, float weight, float* d_accum, cudaStream_t s);
int progressive_accumulate(const bmHit* d_hits1, u64 n, const bmHit* d_hits2, const u32* d_src2, const u32* d_count2, float ambient, float* d_accum,
                           u32* d_covered, cudaStream_t s);
int progressive_resolve(const float* d_accum, const u32* d_covered, u64 n, u32 samples, u32* d_pixels, cudaStream_t s);
int hash_keys(const float* d_points3, u64 n, u32* d_hash, i32* d_cells3, cudaStream_t s);
} // namespace bm
