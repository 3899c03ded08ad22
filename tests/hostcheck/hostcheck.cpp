// Test-only: compiles the product's host+device inline arithmetic (csrc/bm_math.cuh, bm_sat.cuh) for the CPU so that the
// no-GPU test tier can check the very same source against the oracle.  Never shipped, never used by the product.
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../raytracercuda_b200/csrc/bm_sat.cuh"
#include "../../raytracercuda_b200/csrc/bm_grid_host.h"
#include "../../raytracercuda_b200/csrc/bm_wide.cuh"
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

// tri_reject (the conservative pre-test in front of tri_intersect) against tri_intersect itself on n (ray, triangle) pairs:
// counts[0] = accepted by the exact test, [1] = rejected by it, [2] = rejected by the pre-test, [3] = VIOLATIONS: rejected by the pre-test
// although the exact test accepts (must stay 0)
void hc_tri_pretest(const float* o, const float* d, const float* tri9, uint64_t n, uint64_t* counts)
{
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (uint64_t i = 0; i < n; i++)
    {
        V3 a = mk3(tri9[9*i], tri9[9*i+1], tri9[9*i+2]), b = mk3(tri9[9*i+3], tri9[9*i+4], tri9[9*i+5]), c = mk3(tri9[9*i+6], tri9[9*i+7], tri9[9*i+8]);
        V3 oo = mk3(o[3*i], o[3*i+1], o[3*i+2]), dd = mk3(d[3*i], d[3*i+1], d[3*i+2]);
        float u, v;
        const float t = tri_intersect(oo, dd, a, b - a, c - a, u, v);
        const bool accept = !(t == FLT_MAX), rej = tri_reject(oo, dd, a, b - a, c - a);
        counts[accept ? 0 : 1]++;
        if (rej) counts[2]++;
        if (rej && accept) counts[3]++;
    }
}

// ---- the 64-byte wide node (bm_wide.cuh) on the CPU: the product's encoder and traversal step (the very functions the kernels call) over
// a radix tree built here the plain way, so that the no-GPU tier can compare complete traversals with the oracle ----------------------
namespace
{
struct HBox { float lo[3], hi[3]; };
struct HNode { u32 c0, c1, prefix; int plen; HBox box; u32 cell; };
struct HTree
{
    Grid g; HostGrid hg;
    std::vector<u32> leafCode, leafFirst, leafCount, pairTri; std::vector<HBox> leafBox;
    std::vector<HNode> nodes; std::vector<WQ> wide; u32 root = WREF_EMPTY;
    std::vector<V3> v0, e1, e2;
    float leafExt[3];
};
HBox unite(const HBox& a, const HBox& b) { HBox r; for (int k = 0; k < 3; k++) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); } return r; }
const u32 H_LEAF = 0x80000000u;

u32 build_range(HTree& T, int l, int r)
{
    if (l == r) return (u32)l | H_LEAF;
    u32 cl = T.leafCode[l], cr = T.leafCode[r];
    int D = T.g.D;
    int plen = __builtin_clz(cl ^ cr) - (32 - D);
    int bit = D - 1 - plen;                                   // first differing bit (from the LSB)
    int gamma = l;
    while (gamma + 1 <= r && !((T.leafCode[gamma + 1] >> bit) & 1u)) gamma++;
    u32 c0 = build_range(T, l, gamma), c1 = build_range(T, gamma + 1, r);
    HNode& n = T.nodes[gamma];
    n.c0 = c0; n.c1 = c1; n.plen = plen; n.prefix = plen > 0 ? (cl >> (D - plen)) : 0u;
    const HBox& b0 = (c0 & H_LEAF) ? T.leafBox[c0 & ~H_LEAF] : T.nodes[c0].box;
    const HBox& b1 = (c1 & H_LEAF) ? T.leafBox[c1 & ~H_LEAF] : T.nodes[c1].box;
    n.box = unite(b0, b1);
    u32 ix, iy, iz; decode_path(plen, n.prefix, ix, iy, iz);
    n.cell = wide_pack_cell(T.g, ix, iy, iz);
    return (u32)gamma;
}
WSlotBox slot_of(const HBox& b) { WSlotBox s; for (int k = 0; k < 3; k++) { s.lo[k] = b.lo[k]; s.hi[k] = b.hi[k]; } return s; }
}

// traces `nrays` (origin, direction) records through the wide records; mode 0 = reference semantics, 1 = closest t >= 0.
// out_tri / out_t: winning triangle (0xFFFFFFFF = miss) and its t; stats[0] += wide visits, [1] += leaf visits, [2] = slots decoded
// non-conservatively (must stay 0), [3] = records with an unbounded axis.  maxFaces = the reference's per-leaf face cap.
int hc_wide_trace(const float* pos, const uint32_t* idx, uint32_t numTris, const float* wmin, const float* wmax, float minLeaf, uint32_t maxDepth, uint32_t maxFaces,
                  const float* rays, uint64_t nrays, int mode, uint32_t* out_tri, float* out_t, uint64_t* stats)
{
    HTree T;
    if (!build_host_grid(T.hg, wmin, wmax, minLeaf, maxDepth)) return 1;
    fill_grid(T.g, T.hg, T.hg.tab[0].data(), T.hg.tab[1].data(), T.hg.tab[2].data());
    if (getenv("HC_NO_UNIFORM")) T.g.uniform = 0;
    const Grid& g = T.g;
    float ext = 0.f;
    for (int a = 0; a < 3; a++)
    {
        ext = fmaxf(ext, T.hg.wmax[a] - T.hg.wmin[a]);
        float m = 0.f;
        for (size_t i = 0; i + 1 < T.hg.tab[a].size(); i++) m = fmaxf(m, T.hg.tab[a][i + 1] - T.hg.tab[a][i]);
        T.leafExt[a] = m * 1.0001f + (T.hg.wmax[a] - T.hg.wmin[a]) * (1.f / 262144.f);       // = bmTree::view()
    }
    const float pad = ext * (1.f / 1048576.f);
    std::vector<uint64_t> pairs;
    T.v0.resize(numTris); T.e1.resize(numTris); T.e2.resize(numTris);
    for (uint32_t t = 0; t < numTris; t++)
    {
        const float* a = pos + 3 * idx[3 * t]; const float* b = pos + 3 * idx[3 * t + 1]; const float* c = pos + 3 * idx[3 * t + 2];
        V3 va = mk3(a[0], a[1], a[2]), vb = mk3(b[0], b[1], b[2]), vc = mk3(c[0], c[1], c[2]);
        T.v0[t] = va; T.e1[t] = vb - va; T.e2[t] = vc - va;
        Collect col = { &pairs, t };
        descend_triangle(g, va, vb, vc, col);
    }
    std::sort(pairs.begin(), pairs.end());
    for (size_t i = 0; i < pairs.size();)
    {
        size_t j = i; u32 code = (u32)(pairs[i] >> 32);
        HBox bx; for (int k = 0; k < 3; k++) { bx.lo[k] = FLT_MAX; bx.hi[k] = -FLT_MAX; }
        T.leafCode.push_back(code); T.leafFirst.push_back((u32)T.pairTri.size());
        u32 kept = 0;
        while (j < pairs.size() && (u32)(pairs[j] >> 32) == code)
        {
            u32 t = (u32)pairs[j];
            if (kept < maxFaces) { T.pairTri.push_back(t); kept++; }          // BuildTree.cu:53-60: the first maxFaces insertions are kept
            V3 vv[3] = { T.v0[t], T.v0[t] + T.e1[t], T.v0[t] + T.e2[t] };     // boxes enclose every face of the leaf, kept or not (as the product's do)
            for (int q = 0; q < 3; q++) for (int k = 0; k < 3; k++) { float x = get(vv[q], k); bx.lo[k] = fminf(bx.lo[k], x - pad); bx.hi[k] = fmaxf(bx.hi[k], x + pad); }
            j++;
        }
        T.leafCount.push_back(kept); T.leafBox.push_back(bx);
        i = j;
    }
    const int L = (int)T.leafCode.size();
    if (L >= 2)
    {
        T.nodes.resize(L - 1); T.wide.resize(4 * (size_t)(L - 1));
        T.root = build_range(T, 0, L - 1);
        for (int b = 0; b < L - 1; b++)
        {
            const HNode& n = T.nodes[b];
            u32 ref[4] = { n.c0, WREF_EMPTY, n.c1, WREF_EMPTY }, cellAB[2] = { 0, 0 }, lenAB[2] = { WLEN_NONE, WLEN_NONE };
            WSlotBox box[4];
            for (int h = 0; h < 2; h++)
            {
                u32 c = h ? n.c1 : n.c0;
                if (c & H_LEAF) { box[2 * h] = box[2 * h + 1] = slot_of(T.leafBox[c & ~H_LEAF]); continue; }
                const HNode& m = T.nodes[c];
                cellAB[h] = m.cell; lenAB[h] = (u32)m.plen; ref[2 * h] = m.c0; ref[2 * h + 1] = m.c1;
                box[2 * h] = slot_of((m.c0 & H_LEAF) ? T.leafBox[m.c0 & ~H_LEAF] : T.nodes[m.c0].box);
                box[2 * h + 1] = slot_of((m.c1 & H_LEAF) ? T.leafBox[m.c1 & ~H_LEAF] : T.nodes[m.c1].box);
            }
            wide_encode(g, n.cell, (u32)n.plen, cellAB[0], lenAB[0], cellAB[1], lenAB[1], ref, box, &T.wide[4 * (size_t)b]);
            // the decoder must give boxes that enclose the exact ones
            const WQ* q = &T.wide[4 * (size_t)b];
            WCell c0 = wide_cell(g, q[0].x, (int)(q[0].w & 0xFF));
            WFrame f = wide_frame(g, c0.mn, c0.mx, q[3].z, q[3].w, q[0].w >> 24);
            const u32 LW[3] = { q[2].x, q[2].y, q[2].z }, HW[3] = { q[2].w, q[3].x, q[3].y };
            bool unb = false;
            for (int a = 0; a < 3; a++)
            {
                if (f.step[a] == 0.f) unb = true;
                for (int s = 0; s < 4; s++)
                {
                    if (ref[s] == WREF_EMPTY) continue;
                    float lo = wide_fma(wide_magic(LW[a], s), f.step[a], f.addLo[a]), hi = wide_fma(wide_magic(HW[a], s), f.step[a], f.addHi[a]);
                    if (!(lo <= box[s].lo[a] && hi >= box[s].hi[a]) && stats) stats[2]++;
                }
            }
            if (unb && stats) stats[3]++;
        }
    }
    else if (L == 1) T.root = 0u | H_LEAF;
    for (uint64_t i = 0; i < nrays; i++)
    {
        V3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        V3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
        auto cinv = [](float v) { return (v == v) ? fminf(fmaxf(v, -1e30f), 1e30f) : 1e30f; };      // trace.cu cull_inv
        V3 ic = mk3(cinv(inv.x), cinv(inv.y), cinv(inv.z)), oi = o * ic;
        V3 fm = mk3(T.leafExt[0] * fabsf(ic.x), T.leafExt[1] * fabsf(ic.y), T.leafExt[2] * fabsf(ic.z));
        float bt = FLT_MAX; u32 btri = 0xFFFFFFFFu;
        std::vector<u32> stack;
        if (T.root != WREF_EMPTY) stack.push_back(T.root);
        while (!stack.empty())
        {
            u32 cur = stack.back(); stack.pop_back();
            if (!(cur & H_LEAF))
            {
                if (stats) stats[0]++;
                const WQ* q = &T.wide[4 * (size_t)cur];
                u32 push[4];
                int n = mode == 0 ? wide_step<BM_TRACE_REFERENCE>(g, q[0], q[1], q[2], q[3], o, d, inv, ic, oi, fm, bt, push)
                                  : wide_step<BM_TRACE_CLOSEST>(g, q[0], q[1], q[2], q[3], o, d, inv, ic, oi, fm, bt, push);
                for (int k = 0; k < n; k++) stack.push_back(push[k]);
                continue;
            }
            u32 l = cur & ~H_LEAF;
            if (stats) stats[1]++;
            if (mode == 0)
            {
                u32 ix, iy, iz; decode_path(g.D, T.leafCode[l], ix, iy, iz);
                V3 mn, mx; cell_box(g, g.D, ix, iy, iz, mn, mx);
                if (box_ray(mn, mx, o, inv) == FLT_MAX) continue;                       // BuildTree.cu:405-406
            }
            for (u32 k = 0; k < T.leafCount[l]; k++)
            {
                u32 t = T.pairTri[T.leafFirst[l] + k];
                float u, v;
                float dd = tri_intersect(o, d, T.v0[t], T.e1[t], T.e2[t], u, v);
                if (mode != 0 && !(dd >= 0.f)) continue;
                if (dd < bt || (dd == bt && dd != FLT_MAX && t < btri)) { bt = dd; btri = t; }
            }
            if (mode == 0 && bt != FLT_MAX) break;                                      // :427-431
        }
        out_tri[i] = btri; out_t[i] = bt;
    }
    return 0;
}

uint32_t hc_hash3(uint32_t x, uint32_t y, uint32_t z) { return hash3(x, y, z); }
uint32_t hc_hash(uint32_t x) { return hash32(x); }
int32_t hc_map(float f) { return map_cell(f); }
}
