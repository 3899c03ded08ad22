/* beam_b200.h -- C ABI of the B200-native (sm_100a) acceleration-structure build + closest-hit trace path.
 *
 * This is the drop-in boundary for the hot path of d1kkop/RaytracerCuda ("Beam").  In the reference the seam is the set
 * of extern "C" launchers defined in Raytracer/BuildTree.cu / Hash.cu / RTClear.cu and re-declared by hand at the top of
 * each host .cpp (SceneTree.cpp:11-35, SceneHash.cpp:10-24, Camera.cpp:16-19).  Part 1 below exports those symbols with
 * the same names, argument order and calling convention, so the reference's own SceneTree.cpp / Camera.cpp link against
 * libbeam_b200.so unchanged (tests/test_gpu_reference_binaries.py does exactly that).  Note on the convention: glm 0.9.9.0's
 * vec3 has a user-provided copy constructor, so `glm::vec3 bMin` BY VALUE is passed by invisible reference under the Itanium
 * C++ ABI (and every struct > 8 bytes is under Win64): at the binary level by-value vec3, `const vec3&` and `const mat3&` are
 * all pointers, which is how they are declared here.  Part 2 is the explicit, handle-based form the same kernels are reached through when the
 * caller wants hit records, ray batches, build statistics or streams -- things the reference ABI cannot express.
 *
 * All pointers named d_* (and every pointer in Part 1) are DEVICE pointers.  No function falls back to the CPU: if no
 * CUDA device is usable the calls fail with BM_ERROR_CUDA (Part 2) or abort with a message (Part 1, which like the
 * reference returns void, CudaComon.cuh:59-68).
 */
#ifndef BEAM_B200_H
#define BEAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z; } bmVec3;          /* layout == glm::vec3 */
typedef struct { float m[9]; } bmMat3;             /* layout == glm::mat3, column-major: m[3*col + row] */

/* Beam.h:8-16 error codes (returned as uint32 by the host-class mirror), plus codes for Part 2 */
enum {
    BM_OK = 0,                      /* ERROR_ALL_FINE */
    BM_ERROR_NO_VERTICES = 1,
    BM_ERROR_INVALID_PARAMETER = 2,
    BM_ERROR_GPU_ALLOC_FAIL = 3,
    BM_ERROR_INVALID_FORMAT = 4,
    BM_ERROR_RT_CAM_MISMATCH = 5,
    BM_ERROR_UNLOCK_FIRST = 6,
    BM_ERROR_LOCK_FIRST = 7,
    BM_ERROR_NO_RENDER_TARGET = 8,
    BM_ERROR_CUDA = 100,            /* a CUDA runtime call or kernel launch failed; see bmLastErrorString() */
    BM_ERROR_UNSUPPORTED_WORLD = 101, /* world box / leaf size does not give a uniform leaf depth <= 32 */
    BM_ERROR_NOT_BUILT = 102,
    BM_ERROR_CAPACITY = 103
};

/* SharedTypes.h:10-18 StaticMeshData (136 bytes, device pointers). Same layout: bmMarch reads normals (slot 1) through it. */
#define BM_VERTEX_DATA_COUNT 10
typedef struct {
    float*    m_vertexData[BM_VERTEX_DATA_COUNT];
    uint32_t  m_vertexDataSizes[BM_VERTEX_DATA_COUNT];
    uint32_t* m_indices;
    uint32_t  m_numVertices;
    uint32_t  m_numIndices;
} bmStaticMeshData;

/* SharedTypes.h:20-23 bmMaterial {int texture}: one per mesh; texture = index into the texture table, < 0 = none. */
typedef struct { int texture; } bmMaterial;
/* A texture of the table bmShadeHits reads: width*height device texels 0x00RRGGBB, row 0 = top of the image. */
typedef struct { const uint32_t* texels; uint32_t width, height; } bmTexture;

/* Trace2.cu:45-48 bmRay {o, d}: 24 bytes.  Explicit ray batches use this record. */
typedef struct { float o[3]; float d[3]; } bmRay;

/* Hit record written by bmTraceRays / bmTraceCamera: 16 bytes.
 * Replaces bmRayFaceHitResult {bmFace*, dist, u, v} (SharedTypes.h:31-35): the face pointer becomes a global triangle id
 * (mesh base + face index, see bmTreeMeshBases); miss = { FLT_MAX, 0, 0, 0xFFFFFFFF }. */
typedef struct { float t, u, v; uint32_t tri; } bmHit;

/* ------------------------------------------------------------------------------------------------------------------
 * Part 1 -- the reference's own entry points (same symbols).
 * ------------------------------------------------------------------------------------------------------------------ */

/* BuildTree.cu:120-148. `rootNode` identifies the tree (one bmGetNodeSize() buffer per SceneTree, SceneTree.cpp:52).
 * The store / pool pointers are accepted and ignored: node, leaf and triangle storage is owned by the library because
 * the layout changed (sorted leaf ranges + 128-byte nodes instead of pointer-linked pools). Clears the staged mesh list. */
void bmResetScene(void* rootNode, void* faceStore, void* faceGroupStore, void* nodeStore,
                  void* faces, void* facePtrs, void* nodes, uint32_t maxFaces, uint32_t maxFacePtrs, uint32_t maxNodes);

/* BuildTree.cu:259-287. Stages mesh `meshIdx` (device positions vec3[], device indices u32[numIndices]) for the tree
 * identified by treeRootNode, with world box [bMin,bMax].  The sort-based builder needs every mesh before it can emit
 * the hierarchy, so the build itself runs in bmBuildTree, or lazily inside the first bmMarch after an insert. */
void bmInsertMeshInTree(const bmVec3* vertices, const uint32_t* indices, uint32_t numIndices, uint32_t meshIdx,
                        const bmVec3* bMin /* glm::vec3 by value */, const bmVec3* bMax /* glm::vec3 by value */,
                        void* treeRootNode, void* faceStore, void* faceGroupStore, void* nodeStore, void* material);

/* BuildTree.cu:502-520. One primary ray per pixel: dir = orient * initialRays[i] from eye; buffer[i] = shaded pixel
 * ((u32)|n.z*255| << 16 on a hit, 255<<8 on a miss, BuildTree.cu:486-496).  `pitch` is ignored as in the reference. */
void bmMarch(const void* root, const bmVec3* bMin, const bmVec3* bMax, const bmVec3* initialRays,
             uint32_t* buffer, uint32_t pitch, uint32_t width, uint32_t height,
             const bmVec3* eye, const bmMat3* orient, const bmStaticMeshData* meshData);

/* BuildTree.cu:546-555. Kept so SceneTree.cpp:48-59 can size its (now unused) pools. Values equal the reference's. */
uint32_t bmGetFaceSize(void);
uint32_t bmGetFacePtrSize(void);
uint32_t bmGetMaterialSize(void);
uint32_t bmGetNodeSize(void);
uint32_t bmGetFaceStoreSize(void);
uint32_t bmGetNodeStoreSize(void);
uint32_t bmGetFaceGroupStoreSize(void);

/* RTClear.cu:19-31 */
void bmClear(uint32_t* buffer, uint32_t pitch, uint32_t width, uint32_t height, uint32_t clearValue);

/* Hash.cu exports re-declared by SceneHash.cpp:10-24 (TREE_TYPE == HASH, Types.h:10-13; the shipping value is TREE).
 * `cells` identifies the scene.  The reference's hash-grid insertion and march are broken (only the
 * first x-row of a triangle's cell range is inserted, 89 % agreement with the closest hit), so these entry points serve the
 * scene from the same kd-tree as Part 1 over the default world box: bmMarchSpace writes the pixels bmMarch writes.
 * The hash KEY functions are reproduced bit-exactly by bmHashKeys (Part 2). */
void bmResetSpace(void* cells);
void bmInsertMeshInSpace(const bmVec3* vertices, const uint32_t* indices, uint32_t numIndices, void* material, uint32_t meshIdx, void* cells);
void bmMarchSpace(const bmVec3* initialRays, uint32_t* buffer, uint32_t pitch, uint32_t width, uint32_t height,
                  const bmVec3* eye, const bmMat3* orient, const bmStaticMeshData* meshData, void* cells);
uint32_t bmGetCellSize(void);          /* Hash.cu:69-89 sizeof(bmCell) = 8,208 */
uint32_t bmGetMaxHashElements(void);   /* Hash.cu HASH_SIZE = 65,536 */

/* Trace2.cu:188-228 exports re-declared by SceneProgressive.cpp:13-22 (TREE_TYPE == PROGRESSIVE).  bmMarchProgressive has an
 * empty body in the reference and no caller; it is exported as the same no-op (`cb` is the reference's bmHitCallback function
 * pointer).  The size getters return the reference's struct sizes.  The working progressive path is bmCameraRays /
 * bmTraceRays / bmBounceRays / bmProgressiveAccumulate / bmProgressiveResolve below. */
void bmMarchProgressive(void* rays, uint32_t numRays, void* rayBoxQueue, void* leafQueue, void* rayFaceQueue, void* bvhNodes,
                        void* faceClusters, void* hitResultClusters, void* meshDataPtrs, void* cb);
uint32_t bmGetStoreSize(void);
uint32_t bmGetFaceClusterSize(void);
uint32_t bmGetBvhNodeSize(void);
uint32_t bmGetRaySize(void);
uint32_t bmGetRayBoxSize(void);
uint32_t bmGetRayFaceSize(void);
uint32_t bmRayFaceHitClusterSize(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Part 2 -- explicit handle-based entry points (same kernels).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct bmTree bmTree;

/* trace modes */
enum {
    BM_TRACE_REFERENCE = 0,  /* the reference's semantics (BuildTree.cu:394-484): front-to-back by split plane, stop at the
                                first leaf cell that produced a hit, strict-minimum d of any sign inside that leaf, at most
                                256 faces per leaf.  This is the parity mode. */
    BM_TRACE_CLOSEST = 1     /* exhaustive closest hit with t >= 0 (ordered traversal with t-range culling) */
};

const char* bmLastErrorString(void);   /* text of the calling thread's last failed call */
uint64_t bmLaunchCount(void);          /* kernels launched by this library so far in this process */

/* Part 1's launchers return void, and the reference's CUDA_CALL prints and calls exit(1) on failure (CudaComon.cuh:59-68).  A shared
 * library must not end its host: the FIRST failure inside bmInsertMeshInTree / bmMarch / bmClear / ... is kept as a process-wide
 * sticky error (code as in the BM_ERROR_* list, 0 = none) until bmClearStickyError; the failing call has no effect. */
int         bmStickyError(void);
const char* bmStickyErrorString(void);
void        bmClearStickyError(void);
/* The owner of `treeRootNode` is about to free it: drops the tree registered under that pointer together with its device buffers
 * (the reference frees its pools in ~SceneTree, SceneTree.cpp:55-68).  Unknown pointers are ignored. */
void        bmReleaseScene(const void* treeRootNode);
/* Stream (cudaStream_t) used by bmMarch / bmClear / the build of Part-1 trees on the CURRENT device; default 0, the legacy default
 * stream the reference launches on. */
void        bmSetPart1Stream(void* cudaStream);

/* World box, leaf size and depth cap default to SceneTree.cpp:43-45 and BuildTree.cuh:14-18 when null / 0. */
int  bmTreeCreate(bmTree** out, const float* worldMin3, const float* worldMax3, float minLeafSize, uint32_t maxDepth);
void bmTreeDestroy(bmTree* tree);
int  bmTreeSetStream(bmTree* tree, void* cudaStream);          /* cudaStream_t; default 0 */
int  bmTreeReset(bmTree* tree);                                /* forget staged meshes (bmResetScene) */
int  bmTreeAddMesh(bmTree* tree, const float* d_positions, const uint32_t* d_indices, uint32_t numIndices, uint32_t meshIdx);
/* keys -> sort -> leaves -> one bottom-up pass (radix tree + tight boxes) -> node records (see bmTreeNodeLayout). Asynchronous on the tree's stream except for two
 * 8-byte read-backs (pair and leaf counts) that size the later stages. */
int  bmTreeBuild(bmTree* tree);
/* CUDA-event times of the stages of the next builds (costs one synchronise per build): bmTreeLastBuildStages returns
 * "gather <us> classify <us> ... total <us>" for the last timed build ("" before) */
int  bmTreeEnableStageTiming(bmTree* tree, int on);
const char* bmTreeLastBuildStages(const bmTree* tree);
int  bmBuildTree(void* treeRootNode);                          /* same, for a tree made through Part 1 */
bmTree* bmTreeFromRoot(const void* treeRootNode);              /* Part 1 handle -> Part 2 handle (NULL if unknown) */

/* stats[0..7] = triangles, (leaf,triangle) pairs, leaves, internal nodes, leaf depth (code bits), largest leaf,
 *               leaves over 256 faces, meshes */
int  bmTreeStats(const bmTree* tree, uint64_t* stats8);
/* Canonical outputs for parity checks (device -> caller-provided DEVICE buffers, sized from bmTreeStats):
 *   pairs : code[P], tri[P] sorted by (code, global triangle id)        -- "keys" and "sorted primitive order"
 *   nodes : prefixLen[N], prefix[N] of the N = leaves-1 internal nodes   -- "topology"
 *   child boxes: 12 floats per internal node (lo0,hi0,lo1,hi1)           -- "node bounds" */
int  bmTreeCopyPairs(const bmTree* tree, uint32_t* d_code, uint32_t* d_tri);
int  bmTreeCopyTopology(const bmTree* tree, uint32_t* d_prefixLen, uint32_t* d_prefix, uint32_t* d_child0, uint32_t* d_child1);
int  bmTreeCopyChildBoxes(const bmTree* tree, float* d_boxes12);
int  bmTreeMeshBases(const bmTree* tree, uint32_t* h_bases, uint32_t capacity);   /* host array: first global id of each mesh */

/* Replication (multi-GPU: build once, copy to peers).  The traversal arrays hold indices only, so they can be copied byte for
 * byte.  sizes8 = bytes of {nodes, leaves, triangle records, cluster boxes, soup} then {leaves, pairs, triangles} counts.
 * bmTreeExport copies the arrays into caller-owned device buffers (on the tree's stream); after those bytes have been moved to
 * another GPU (ncclBroadcast, P2P ...) bmTreeImport wraps them as a read-only tree there: no copy, the caller keeps the buffers
 * alive; such a tree can be traced (and feed bmBounceRays) but not rebuilt or inspected with bmTreeCopy*. */
int  bmTreeArraySizes(const bmTree* tree, uint64_t* sizes8);
int  bmTreeExport(const bmTree* tree, void* d_nodes, void* d_leaves, void* d_ltris, void* d_clusters, void* d_soup);
int  bmTreeRoot(const bmTree* tree, uint32_t* root);           /* reference of the root: node index, or 0x80000000 | 0 when the only leaf is the root */
int  bmTreeImport(bmTree** out, const void* d_nodes, const void* d_leaves, const void* d_ltris, const void* d_clusters, const void* d_soup,
                  uint64_t numLeaves, uint64_t numPairs, uint64_t numTris, uint32_t root, uint32_t nodeLayout,
                  const float* worldMin3, const float* worldMax3, float minLeafSize, uint32_t maxDepth);   /* layout + world of the source: bmTreeNodeLayout, bmTreeWorld */
/* Node records of a built tree.  BM_NODES_BINARY: 128 bytes per internal node of the binary radix tree (two child boxes, child references,
 * the reference node's cell box / split plane / axis precomputed); chosen while the node array stays cache-resident, where traversal is
 * issue-bound.  BM_NODES_WIDE: 64 bytes, 64-byte aligned, per internal node: the node and its two children collapsed into a 4-wide record --
 * integer cell coordinates of the three kd nodes (their boxes and planes are rebuilt exactly in the kernel) + four slot references + four
 * culling boxes quantised to 8 bits per coordinate, rounded outwards (raytracercuda_b200/csrc/bm_wide.cuh); chosen for trees beyond the
 * L2, where traversal is bound by memory round trips and bytes.  BM_NODE_LAYOUT=binary|wide|both and BM_WIDE_MIN_MB override the choice.
 * Results do not depend on the layout. */
enum { BM_NODES_BINARY = 0, BM_NODES_WIDE = 1, BM_NODES_BOTH = 2 };
int  bmTreeNodeLayout(const bmTree* tree, uint32_t* layout);
int  bmTreeWorld(const bmTree* tree, float* worldMin3, float* worldMax3, float* minLeafSize, uint32_t* maxDepth);   /* any pointer may be null */
/* Parity aid for the wide records: the decoded culling boxes (24 floats per node: lo.xyz, hi.xyz of slots 0..3) and the slot references
 * (4 per node; bit 31 = leaf, 0x7FFFFFFF = empty) into caller-provided DEVICE buffers. */
int  bmTreeCopyWideSlots(const bmTree* tree, float* d_boxes24, uint32_t* d_refs4);
/* A second, read-only handle on the arrays of a built tree (no copy) with its own stream (bmTreeSetStream), work counter and bounce
 * scratch: launches issued through different views of one tree may run concurrently, e.g. the two halves of a frame on two streams so
 * that the tail of one half's traversal is filled by the other half's work.  Valid until `tree` is rebuilt or destroyed. */
int  bmTreeView(const bmTree* tree, bmTree** out);

/* Explicit ray batch -> hit records. d_counters (nullable, FOUR uint64) accumulates {node + leaf visits, triangle tests, cluster
 * boxes tested} (the traversal work behind the bytes-per-ray roofline model) and keeps in [3] the largest number of such steps any
 * single ray needed (the launch's critical path). */
int  bmTraceRays(bmTree* tree, const bmRay* d_rays, uint64_t numRays, int mode, bmHit* d_hits, uint64_t* d_counters);
/* How bmTraceRays walks trees far larger than the L2.  The top levels of the reference's kd-tree are an implicit grid (midpoint splits of the
 * world box, BuildTree.cu:196-216), so a batch can be cut into one SEGMENT per (ray, brick of that grid entered), the segments sorted by
 * brick and traced brick by brick out of the cache; per-ray results are merged in the order the reference's own traversal would have met
 * the bricks (raytracercuda_b200/csrc/bricks.cu).  Hits are identical to the direct traversal in both modes.  mode: -1 automatic
 * (traversal arrays above BM_BRICK_MIN_MB = 512 MB and >= 2^18 rays; BM_BRICKS=0|1 overrides), 0 never, 1 always; bits: depth of the
 * brick grid, 3..15, 0 = default (6 = 4 x 4 x 4 bricks, fewer for trees below 256 MB; BM_BRICK_BITS overrides).  A bricked call synchronises the tree's stream once (the
 * segment count sizes the sort). */
int  bmTreeSetBricks(bmTree* tree, int mode, int bits);
/* Same, with the ray count read from DEVICE memory when the kernel starts (*d_numRays <= maxRays): lets a producer such as
 * bmBounceRays feed the trace without a host round trip. */
int  bmTraceRaysIndirect(bmTree* tree, const bmRay* d_rays, const uint32_t* d_numRays, uint64_t maxRays, int mode, bmHit* d_hits, uint64_t* d_counters);
/* A ray batch that IS an image -- width x (numRays / width) rays in row-major pixel order, e.g. the jittered camera rays of one sample of
 * a multi-sample frame (bmCameraRays) -- traced with the camera launches' scheduling: 8x4 pixel tiles per warp and the temporal tile
 * schedule (consecutive samples of a frame share it).  Same hits as bmTraceRays. */
int  bmTraceRaysTiled(bmTree* tree, const bmRay* d_rays, uint64_t numRays, uint32_t width, int mode, bmHit* d_hits, uint64_t* d_counters);
/* Camera rays (dir = orient * initialRays[i] from eye), hit records and/or shaded pixels (either may be null).
 * width = row length of the image the table describes (numRays % width == 0): warps then start on 8x4 pixel tiles; 0 = trace in
 * linear order.  Results are always written at index i of the table, whatever the order of traversal. */
int  bmTraceCamera(bmTree* tree, const float* d_initialRays3, uint64_t numRays, uint32_t width, const float* eye3, const float* orient9,
                   int mode, bmHit* d_hits, uint32_t* d_pixels, const bmStaticMeshData* d_meshData, uint64_t* d_counters);

/* bmTraceCamera with the sample jitter applied where the ray is formed (in the trace kernel's registers): sample == 0xFFFFFFFF is the
 * exact pixel-centre table entry, otherwise the direction is jittered inside the pixel exactly as bmCameraRays would write it -- a
 * multi-sample frame needs no per-sample ray buffer (SceneProgressive.cpp:27-77 is the stub this stands in for). */
int  bmTraceCameraJittered(bmTree* tree, const float* d_initialRays3, uint64_t numRays, uint32_t width, const float* eye3, const float* orient9,
                           float jitterX, float jitterY, uint32_t sample, uint32_t seed, int mode,
                           bmHit* d_hits, uint32_t* d_pixels, const bmStaticMeshData* d_meshData, uint64_t* d_counters);

/* Ray-batch producers for the multi-sample / bounce workloads (defined by this repo; the reference has no secondary rays,
 * SURVEY.md section 8d).  bmCameraRays writes the camera's rays as bmRay records; sample == 0xFFFFFFFF means the exact
 * pixel-centre table entry (identical to what bmMarch traces), otherwise the direction is jittered inside the pixel
 * (jitterX/Y = pixel size in units of zoom) by a counter-based hash of (ray, sample, seed).
 * bmBounceRays writes one cosine-weighted diffuse bounce ray per hit (origin = hit + 1e-4 n, geometric normal n facing the
 * incoming ray), compacted in ray order; d_srcIndex (nullable) receives the index of the primary ray of each bounce ray and
 * *d_count (device uint32) the number of bounce rays.  d_out / d_srcIndex need room for numRays entries. */
int  bmCameraRays(const float* d_initialRays3, uint64_t numRays, uint32_t width, const float* eye3, const float* orient9,
                  float jitterX, float jitterY, uint32_t sample, uint32_t seed, bmRay* d_out, void* cudaStream);
int  bmBounceRays(bmTree* tree, const bmRay* d_rays, const bmHit* d_hits, uint64_t numRays, uint32_t sample, uint32_t bounce,
                  uint32_t seed, bmRay* d_out, uint32_t* d_srcIndex, uint32_t* d_count);
/* bmBounceRays for the hits of a bmTraceCamera / bmMarch launch: the primary rays are re-formed from the camera table */
int  bmBounceRaysCamera(bmTree* tree, const float* d_initialRays3, const float* eye3, const float* orient9, const bmHit* d_hits,
                        uint64_t numRays, uint32_t sample, uint32_t bounce, uint32_t seed, bmRay* d_out, uint32_t* d_srcIndex, uint32_t* d_count);

/* ... of a bmTraceCameraJittered launch (cameraSample / cameraSeed / jitter as given there) */
int  bmBounceRaysCameraJittered(bmTree* tree, const float* d_initialRays3, const float* eye3, const float* orient9, float jitterX, float jitterY,
                                uint32_t cameraSample, uint32_t cameraSeed, const bmHit* d_hits, uint64_t numRays,
                                uint32_t sample, uint32_t bounce, uint32_t seed, bmRay* d_out, uint32_t* d_srcIndex, uint32_t* d_count);

/* Append form of bmBounceRaysCameraJittered: the rays go BEHIND the *d_total rays already in d_out / d_srcIndex (device uint32, in/out; the
 * caller zeroes it before the first call and sizes the buffers for all samples it will append), so that the bounce rays of several samples
 * of a frame are traced by one bmTraceRaysIndirect launch -- one sample of a small scene yields too few to fill the GPU.  d_srcIndex
 * receives the pixel (index in the camera table) of every ray; the same pixel may appear once per appended sample. */
int  bmBounceRaysCameraAppend(bmTree* tree, const float* d_initialRays3, const float* eye3, const float* orient9, float jitterX, float jitterY,
                              uint32_t cameraSample, uint32_t cameraSeed, const bmHit* d_hits, uint64_t numRays,
                              uint32_t sample, uint32_t bounce, uint32_t seed, bmRay* d_out, uint32_t* d_srcIndex, uint32_t* d_total);

/* Multi-sample camera launch: `sampleCount` jittered samples (firstSample .. firstSample + sampleCount - 1, each the ray
 * bmTraceCameraJittered(..., sample, seed, ...) traces) of every pixel in ONE launch; d_hits holds sampleCount * numRays records,
 * sample-major (record of sample k, pixel i at k * numRays + i).  The launch walks the image in 8x4 tiles, the samples of a tile one after
 * the other, expensive tiles first (temporal tile schedule): a multi-sample frame no longer pays one launch tail per sample. */
int  bmTraceCameraSamples(bmTree* tree, const float* d_initialRays3, uint64_t numRays, uint32_t width, const float* eye3, const float* orient9,
                          float jitterX, float jitterY, uint32_t firstSample, uint32_t sampleCount, uint32_t seed, int mode, bmHit* d_hits, uint64_t* d_counters);

/* Bounce launch: every pixel whose primary ray hit (d_hits1, laid out as bmTraceCameraSamples / bmTraceCamera[Jittered] wrote it) traces the
 * diffuse bounce ray of that hit -- bit for bit the ray bmBounceRaysCamera[Jittered] would append to a list, formed in registers instead --
 * as a closest hit with t >= 0.  No compaction and no ray list: the rays of a warp are the hits of one 8x4 pixel tile, and the launch runs
 * under its own temporal tile schedule.
 *   cameraSample0  0xFFFFFFFF: the primary rays were the exact pixel-centre rays (sampleCount must be 1); else the first camera sample
 *   bounceSample0  index of the launch's first sample in the bounce rays' random sequence (bmBounceRays' `sample`)
 *   d_hits2        nullable; sampleCount * numRays records, written ONLY where the primary ray hit
 *   d_accum, d_covered  nullable; the progressive estimator (bmProgressiveAccumulate) applied where a bounce ray ends:
 *                  d_accum[i] += ambient + (escaped ? weight : 0), d_covered[i] = 1
 *   d_rayCount     nullable, device uint64: += number of bounce rays traced
 * A camera launch that writes hit records also flags the 8x4 tiles in which a ray hit.  When d_hits1 is the buffer the LAST camera launch
 * on this handle wrote (same pointer, numRays, width and sampleCount) the bounce launch leaves the unflagged tiles out of its work list;
 * the caller must not have changed the records in between (trace into another buffer instead).  Any other camera launch on the handle,
 * a rebuild, or BM_BOUNCE_LIST=0 makes the bounce launch walk every tile.  Results are the same either way. */
int  bmTraceBounceCamera(bmTree* tree, const float* d_initialRays3, uint64_t numRays, uint32_t width, const float* eye3, const float* orient9,
                         float jitterX, float jitterY, uint32_t cameraSample0, uint32_t sampleCount, uint32_t cameraSeed, const bmHit* d_hits1,
                         uint32_t bounceSample0, uint32_t bounce, uint32_t seed, bmHit* d_hits2, float ambient, float weight, float* d_accum,
                         uint32_t* d_covered, uint64_t* d_rayCount, uint64_t* d_counters);

/* Progressive accumulation over samples (defined by this repo; the reference's SceneProgressive is an empty stub): one sample of a
 * pixel = `ambient` if its primary ray hit + (1 - ambient) if the diffuse bounce ray of that hit escaped.  d_accum (float) and
 * d_covered (uint32) are per pixel and must start zeroed; d_hits2 / d_src2 / d_count2 are bmBounceRays' outputs after tracing (pass
 * d_hits2 = NULL for primary only).  bmProgressiveResolve writes grey = accum / samples into the pixel's three colour bytes, and the
 * reference's miss colour (255 << 8) where no sample ever hit. */
int  bmProgressiveAccumulate(const bmHit* d_hits1, uint64_t numRays, const bmHit* d_hits2, const uint32_t* d_src2, const uint32_t* d_count2,
                             float ambient, float* d_accum, uint32_t* d_covered, void* cudaStream);
/* the bounce half of bmProgressiveAccumulate for a list that holds several samples (bmBounceRaysCameraAppend): d_accum[d_src2[k]] +=
 * weight for every escaped ray k < *d_count2, with atomic adds (a pixel may occur once per sample); maxRays sizes the launch */
int  bmProgressiveAccumulateBounce(const bmHit* d_hits2, const uint32_t* d_src2, const uint32_t* d_count2, uint64_t maxRays, float weight, float* d_accum, void* cudaStream);
int  bmProgressiveResolve(const float* d_accum, const uint32_t* d_covered, uint64_t numRays, uint32_t samples, uint32_t* d_pixels, void* cudaStream);

/* Deferred shading of hit records (what the reference's bmHitCallback / bmMaterial / VERTEX_DATA_UV1 declare but never implement,
 * SharedTypes.h:20-37, Beam.h:21): miss -> 255<<8; hit on a mesh without texture or UV slot -> the reference's pixel
 * (u32)|n.z*255| << 16 (BuildTree.cu:486-496, bit for bit what bmTraceCamera writes); hit on a textured mesh (d_materials[mesh].texture
 * in [0, numTextures), UV in vertex-data slot 2 with >= 2 components) -> nearest texel at the interpolated uv (repeat wrap, v up),
 * every channel scaled by |n.z|, packed 0x00RRGGBB.  d_materials may be null (no mesh textured).  Needs the tree for the
 * triangle id -> (mesh, face) mapping; runs on the tree's stream. */
int  bmShadeHits(bmTree* tree, const bmHit* d_hits, uint64_t numRays, const bmStaticMeshData* d_meshData, const bmMaterial* d_materials,
                 const bmTexture* d_textures, uint32_t numTextures, uint32_t* d_pixels);

/* Spatial-hash key functions of Hash.cu:17-64 evaluated on the device (bit-exact targets):
 * out[i] = bmHash3(make_uint3(bmMap(x), bmMap(y), bmMap(z))) and the three bmMap cell coordinates. */
int  bmHashKeys(const float* d_points3, uint64_t n, uint32_t* d_hash, int32_t* d_cells3, void* cudaStream);

#ifdef __cplusplus
}
#endif
#endif /* BEAM_B200_H */
