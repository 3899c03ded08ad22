// rays.cu -- ray-batch producers for the workloads the reference only names (SURVEY.md section 8d "Secondary / spp rays"):
//   * camera_rays: materialise bmRay{o,d} records from the camera's direction table (optionally jittered inside the pixel
//     for sample s of a multi-sample frame);
//   * bounce_rays: one diffuse bounce ray per primary hit: origin = hit + 1e-4*n, direction cosine-weighted about the
//     geometric normal n (flipped towards the incoming ray's origin), RNG = counter-based hash of (ray index, sample,
//     bounce, seed).  Hits are compacted in ray order (flag -> exclusive scan -> scatter), so the output order is
//     deterministic; srcIndex maps each bounce ray back to its primary ray.
// Parity for these rays is judged on identical ray buffers (the oracle traces the buffer the GPU produced), so the
// sin/cos/sqrt used here need not be bit-compatible with any host implementation.
#include "bm_device.cuh"

namespace bm
{
namespace
{
__global__ void camera_rays_kernel(CameraIn cam, u64 n, bmRay* __restrict__ out)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 o, d;
    camera_ray(cam, i, o, d);
    bmRay r;
    r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
    out[i] = r;
}

__global__ void add_count_kernel(const u32* __restrict__ count, u64 add, u64* total) { *total += (count ? (u64)*count : 0ull) + add; }

template <bool CAMERA>
__global__ void bounce_kernel(const bmRay* __restrict__ rays, CameraIn cam, const bmHit* __restrict__ hits, const u32* __restrict__ offsets, u64 n,
                              const float4* __restrict__ soup, u32 sample, u32 bounce, u32 seed, bmRay* __restrict__ out, u32* __restrict__ srcIndex,
                              const u32* __restrict__ appendBase)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 hf = reinterpret_cast<const float4*>(hits)[i];
    u32 tri = __float_as_uint(hf.w);
    if (tri == 0xFFFFFFFFu) return;
    bmRay r;
    if (CAMERA)
    {
        // the primary ray exactly as bmMarch / bmTraceCamera formed it (camera_ray: same function, same bits)
        V3 co, cd;
        camera_ray(cam, i, co, cd);
        r.o[0] = co.x; r.o[1] = co.y; r.o[2] = co.z; r.d[0] = cd.x; r.d[1] = cd.y; r.d[2] = cd.z;
    }
    else r = rays[i];
    V3 bo, bd;
    bounce_ray(mk3(r.o[0], r.o[1], r.o[2]), mk3(r.d[0], r.d[1], r.d[2]), hf.x, tri, soup, i, sample, bounce, seed, bo, bd);
    bmRay o;
    o.o[0] = bo.x; o.o[1] = bo.y; o.o[2] = bo.z; o.d[0] = bd.x; o.d[1] = bd.y; o.d[2] = bd.z;
    u32 k = offsets[i] + (appendBase ? *appendBase : 0u);      // append mode: behind the rays other launches already put into the list
    out[k] = o;
    if (srcIndex) srcIndex[k] = (u32)i;
}
// ---- progressive accumulation (the estimator this repo defines for the multi-sample configs; the reference's SceneProgressive is an
// empty stub, SceneProgressive.cpp:27-77).  One sample of a pixel = ambient if the primary ray hit, plus (1 - ambient) if the diffuse
// bounce ray of that hit escaped (cosine-weighted ambient occlusion); pixels whose primary rays never hit keep the reference's miss
// colour.  accum is one float per pixel, summed over samples; resolve divides by the sample count and packs grey into the red, green
// and blue bytes. ----------------------------------------------------------------------------------------------------------------
__global__ void accumulate_primary_kernel(const bmHit* __restrict__ hits, u64 n, float ambient, float* __restrict__ accum, u32* __restrict__ covered)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (hits[i].tri != 0xFFFFFFFFu) { accum[i] += ambient; covered[i] = 1u; }
}
__global__ void accumulate_bounce_kernel(const bmHit* __restrict__ hits2, const u32* __restrict__ src, const u32* __restrict__ count, float weight,
                                         float* __restrict__ accum)
{
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *count) return;
    if (hits2[k].tri == 0xFFFFFFFFu) accum[src[k]] += weight;       // src is injective (one bounce ray per primary hit): no atomics needed
}
// a list that holds the bounce rays of SEVERAL samples may name a pixel more than once: atomic adds
__global__ void accumulate_bounce_atomic_kernel(const bmHit* __restrict__ hits2, const u32* __restrict__ src, const u32* __restrict__ count, float weight,
                                                float* __restrict__ accum)
{
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *count) return;
    if (hits2[k].tri == 0xFFFFFFFFu) atomicAdd(&accum[src[k]], weight);
}
__global__ void append_count_kernel(u32* total, const u32* __restrict__ count) { *total += *count; }
__global__ void resolve_kernel(const float* __restrict__ accum, const u32* __restrict__ covered, u64 n, float invSamples, u32* __restrict__ pixels)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!covered[i]) { pixels[i] = 255u << 8; return; }
    float v = fminf(fmaxf(accum[i] * invSamples, 0.f), 1.f);
    u32 g = (u32)(v * 255.f);
    pixels[i] = (g << 16) | (g << 8) | g;
}
} // namespace

int progressive_accumulate(const bmHit* d_hits1, u64 n, const bmHit* d_hits2, const u32* d_src2, const u32* d_count2, float ambient, float* d_accum,
                           u32* d_covered, cudaStream_t s)
{
    if (!n) return BM_OK;
    accumulate_primary_kernel<<<div_up(n, 256), 256, 0, s>>>(d_hits1, n, ambient, d_accum, d_covered);
    BM_LAUNCHED();
    if (d_hits2)
    {
        accumulate_bounce_kernel<<<div_up(n, 256), 256, 0, s>>>(d_hits2, d_src2, d_count2, 1.f - ambient, d_accum);
        BM_LAUNCHED();
    }
    return BM_OK;
}
int progressive_accumulate_bounce(const bmHit* d_hits2, const u32* d_src2, const u32* d_count2, u64 maxRays, float weight, float* d_accum, cudaStream_t s)
{
    if (!maxRays) return BM_OK;
    accumulate_bounce_atomic_kernel<<<div_up(maxRays, 256), 256, 0, s>>>(d_hits2, d_src2, d_count2, weight, d_accum);
    BM_LAUNCHED();
    return BM_OK;
}
int progressive_resolve(const float* d_accum, const u32* d_covered, u64 n, u32 samples, u32* d_pixels, cudaStream_t s)
{
    if (!n) return BM_OK;
    resolve_kernel<<<div_up(n, 256), 256, 0, s>>>(d_accum, d_covered, n, 1.f / (float)(samples ? samples : 1), d_pixels);
    BM_LAUNCHED();
    return BM_OK;
}

int camera_rays(const CameraIn& cam, u64 n, bmRay* d_out, cudaStream_t s)
{
    if (!n) return BM_OK;
    camera_rays_kernel<<<div_up(n, 256), 256, 0, s>>>(cam, n, d_out);
    BM_LAUNCHED();
    return BM_OK;
}

int add_count(const u32* d_count, u64 add, u64* d_total, cudaStream_t s)
{
    add_count_kernel<<<1, 1, 0, s>>>(d_count, add, d_total);
    BM_LAUNCHED();
    return BM_OK;
}

int bounce_rays(const bmRay* d_rays, const CameraIn* cam, const bmHit* d_hits, u64 n, const float4* d_soup, u32 sample, u32 bounce, u32 seed,
                u32* d_flagsAndOffsets, void* d_scanTemp, bmRay* d_out, u32* d_srcIndex, u32* d_count, cudaStream_t s, u32* d_appendTotal)
{
    // d_appendTotal (nullable, device): append mode -- the rays go behind the *d_appendTotal rays already in d_out / d_srcIndex and the
    // total grows by this launch's count (d_count then receives this launch's own count)
    if (!n) { BM_CUDA(cudaMemsetAsync(d_count, 0, 4, s)); return BM_OK; }
    BM_TRY(exclusive_scan_hit_flags(d_hits, d_flagsAndOffsets, n, d_count, d_scanTemp, s));
    if (cam) bounce_kernel<true><<<div_up(n, 256), 256, 0, s>>>(nullptr, *cam, d_hits, d_flagsAndOffsets, n, d_soup, sample, bounce, seed, d_out, d_srcIndex, d_appendTotal);
    else { CameraIn none = make_camera(nullptr, nullptr, nullptr); bounce_kernel<false><<<div_up(n, 256), 256, 0, s>>>(d_rays, none, d_hits, d_flagsAndOffsets, n, d_soup, sample, bounce, seed, d_out, d_srcIndex, d_appendTotal); }
    BM_LAUNCHED();
    if (d_appendTotal) { append_count_kernel<<<1, 1, 0, s>>>(d_appendTotal, d_count); BM_LAUNCHED(); }
    return BM_OK;
}
} // namespace bm
