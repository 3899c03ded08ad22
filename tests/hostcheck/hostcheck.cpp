// Test-only: compiles the product's host+device inline arithmetic (csrc/bm_math.cuh, bm_sat.cuh) for the CPU so that the
// no-GPU test tier can check the very same source against the oracle.  Never shipped, never used by the product.
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../raytracercuda_b200/csrc/bm_sat.cuh"
#include "../../raytracercuda_b200/csrc/bm_grid_host.h"
using namespace bm;

struct Collect { std::vector<uint64_t>* out; uint32_t tri; void operator()(u32 code, u32, u32, u32) const { out->push_back(((uint64_t)code << 32) | tri); } };

extern "C"
{
// pairs for an indexed mesh with the default world; returns count, fills sorted (code, tri) when out pointers non-null
uint64_t hc_pairs(const float* pos, const uint32_t* idx, uint32_t numTris, const float* wmin, const float* wmax, float minLeaf, uint32_t maxDepth,
                  uint32_t* code, uint32_t* tri, uint64_t cap, int* D)
{
    HostGrid hg;
    if (!build_host_grid(hg, wmin, wmax, minLeaf, maxDepth)) return ~0ull;
    Grid g; fill_grid(g, hg, hg.tab[0].data(), hg.tab[1].data(), hg.tab[2].data());
    if (getenv("HC_NO_UNIFORM")) g.uniform = 0;
    if (D) *D = hg.D;
    std::vector<uint64_t> out;
    for (uint32_t t = 0; t < numTris; t++)
    {
        const float* a = pos + 3 * idx[3 * t]; const float* b = pos + 3 * idx[3 * t + 1]; const float* c = pos + 3 * idx[3 * t + 2];
        Collect col = { &out, t };
        descend_triangle(g, mk3(a[0], a[1], a[2]), mk3(b[0], b[1], b[2]), mk3(c[0], c[1], c[2]), col);
    }
    std::sort(out.begin(), out.end());
    if (code && tri) for (size_t i = 0; i < out.size() && i < cap; i++) { code[i] = (uint32_t)(out[i] >> 32); tri[i] = (uint32_t)out[i]; }
    return out.size();
}
// the flat path of key generation (classify_kernel's logic): exact per-axis ranges, pair_test per candidate cell when there are
// at most 64 candidates, descent otherwise
uint64_t hc_pairs_flat(const float* pos, const uint32_t* idx, uint32_t numTris, const float* wmin, const float* wmax, float minLeaf, uint32_t maxDepth,
                       uint32_t* code, uint32_t* tri, uint64_t cap, uint64_t* numFlatCandidates)
{
    HostGrid hg;
    if (!build_host_grid(hg, wmin, wmax, minLeaf, maxDepth)) return ~0ull;
    Grid g; fill_grid(g, hg, hg.tab[0].data(), hg.tab[1].data(), hg.tab[2].data());
    if (getenv("HC_NO_UNIFORM")) g.uniform = 0;
    float ext = 0.f; for (int a = 0; a < 3; a++) ext = std::max(ext, wmax[a] - wmin[a]);
    float margin = ext * (1.f / 524288.f);
    std::vector<uint64_t> out;
    uint64_t cand = 0;
    for (uint32_t t = 0; t < numTris; t++)
    {
        const float* a = pos + 3 * idx[3 * t]; const float* b = pos + 3 * idx[3 * t + 1]; const float* c = pos + 3 * idx[3 * t + 2];
        V3 va = mk3(a[0], a[1], a[2]), vb = mk3(b[0], b[1], b[2]), vc = mk3(c[0], c[1], c[2]);
        int lo[3], hi[3];
        axis_range(g, 0, fminf(a[0], fminf(b[0], c[0])), fmaxf(a[0], fmaxf(b[0], c[0])), lo[0], hi[0]);
        axis_range(g, 1, fminf(a[1], fminf(b[1], c[1])), fmaxf(a[1], fmaxf(b[1], c[1])), lo[1], hi[1]);
        axis_range(g, 2, fminf(a[2], fminf(b[2], c[2])), fmaxf(a[2], fmaxf(b[2], c[2])), lo[2], hi[2]);
        if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) continue;
        uint64_t n = (uint64_t)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
        if (n > 64) { Collect col = { &out, t }; descend_triangle(g, va, vb, vc, col); continue; }
        cand += n;
        TriSat ts; trisat_init(ts, va, vb, vc);
        for (int iz = lo[2]; iz <= hi[2]; iz++) for (int iy = lo[1]; iy <= hi[1]; iy++) for (int ix = lo[0]; ix <= hi[0]; ix++)
            if (pair_test(g, ts, (u32)ix, (u32)iy, (u32)iz, margin)) out.push_back(((uint64_t)encode_path(g.D, ix, iy, iz) << 32) | t);
    }
    std::sort(out.begin(), out.end());
    if (code && tri) for (size_t i = 0; i < out.size() && i < cap; i++) { code[i] = (uint32_t)(out[i] >> 32); tri[i] = (uint32_t)out[i]; }
    if (numFlatCandidates) *numFlatCandidates = cand;
    return out.size();
}
int hc_grid(const float* wmin, const float* wmax, float minLeaf, uint32_t maxDepth, int* k3, int* D, float* tabx, float* taby, float* tabz)
{
    HostGrid hg;
    if (!build_host_grid(hg, wmin, wmax, minLeaf, maxDepth)) return 0;
    for (int a = 0; a < 3; a++) k3[a] = hg.k[a]; *D = hg.D;
    if (tabx) memcpy(tabx, hg.tab[0].data(), hg.tab[0].size() * 4);
    if (taby) memcpy(taby, hg.tab[1].data(), hg.tab[1].size() * 4);
    if (tabz) memcpy(tabz, hg.tab[2].data(), hg.tab[2].size() * 4);
    return 1;
}
uint32_t hc_encode(int depth, uint32_t ix, uint32_t iy, uint32_t iz) { return encode_path(depth, ix, iy, iz); }
void hc_decode(int depth, uint32_t p, uint32_t* xyz) { decode_path(depth, p, xyz[0], xyz[1], xyz[2]); }
float hc_box_ray(const float* mn, const float* mx, const float* o, const float* id) { return box_ray(mk3(mn[0],mn[1],mn[2]), mk3(mx[0],mx[1],mx[2]), mk3(o[0],o[1],o[2]), mk3(id[0],id[1],id[2])); }
float hc_tri_intersect(const float* o, const float* d, const float* v0, const float* v1, const float* v2, float* u, float* v)
{
    V3 a = mk3(v0[0],v0[1],v0[2]), b = mk3(v1[0],v1[1],v1[2]), c = mk3(v2[0],v2[1],v2[2]);
    return tri_intersect(mk3(o[0],o[1],o[2]), mk3(d[0],d[1],d[2]), a, b - a, c - a, *u, *v);
}
uint32_t hc_hash3(uint32_t x, uint32_t y, uint32_t z) { return hash3(x, y, z); }
uint32_t hc_hash(uint32_t x) { return hash32(x); }
int32_t hc_map(float f) { return map_cell(f); }
}
