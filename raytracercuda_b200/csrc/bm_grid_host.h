// bm_grid_host.h -- host-side construction of the implicit grid tables (see bm_math.cuh, struct Grid).
#pragma once
#include <vector>
#include <algorithm>
#include "bm_math.cuh"

namespace bm
{
struct HostGrid
{
    float wmin[3], wmax[3];
    float minLeaf; uint32_t maxDepth;
    int k[3]; int D;
    std::vector<float> tab[3];
    int uniform; float w[3];
};

// fills the device-facing Grid from a HostGrid (table pointers supplied by the caller)
inline void fill_grid(Grid& g, const HostGrid& h, const float* t0, const float* t1, const float* t2)
{
    g.tab[0] = t0; g.tab[1] = t1; g.tab[2] = t2;
    for (int a = 0; a < 3; a++) { g.k[a] = h.k[a]; g.wmin[a] = h.wmin[a]; g.w[a] = h.w[a]; g.wext[a] = h.wmax[a] - h.wmin[a]; }
    g.D = h.D; g.uniform = h.uniform; g.subBits = 9;
}

// Replays the reference's subdivision rule on the world box (BuildTree.cu:196-216): at depth d the node is a leaf when
// min(edge) < minLeaf or d == maxDepth-1, else it is halved on axis d%3 at s = .5f*(min+max).  All nodes of one depth must
// take the same decision (true for the reference's own +-30 box, where every leaf is at depth 31) and the leaf depth must
// fit a 32-bit path code; otherwise the world is rejected.  Returns false in that case.
inline bool build_host_grid(HostGrid& g, const float wmin[3], const float wmax[3], float minLeaf, uint32_t maxDepth)
{
    for (int a = 0; a < 3; a++) { g.wmin[a] = wmin[a]; g.wmax[a] = wmax[a]; g.tab[a].assign({ wmin[a], wmax[a] }); g.k[a] = 0; }
    g.minLeaf = minLeaf; g.maxDepth = maxDepth;
    for (int a = 0; a < 3; a++) if (!(wmax[a] > wmin[a])) return false;
    for (uint32_t d = 0;; d++)
    {
        // leaf decision: uniform iff on every axis all intervals agree about "< minLeaf"
        bool anyAxisAllSmall = false, mixed = false;
        for (int a = 0; a < 3; a++)
        {
            bool allSmall = true, noneSmall = true;
            for (size_t i = 0; i + 1 < g.tab[a].size(); i++)
            {
                float w = g.tab[a][i + 1] - g.tab[a][i];
                if (!(w > 0.f)) return false;
                if (w < minLeaf) noneSmall = false; else allSmall = false;
            }
            if (allSmall) anyAxisAllSmall = true; else if (!noneSmall) mixed = true;
        }
        if (anyAxisAllSmall || d == maxDepth - 1) { g.D = (int)d; break; }
        if (mixed) return false;
        if (d >= 32) return false;
        int a = d % 3;
        std::vector<float> nt(2 * g.tab[a].size() - 1);
        for (size_t i = 0; i + 1 < g.tab[a].size(); i++)
        {
            nt[2 * i] = g.tab[a][i];
            nt[2 * i + 1] = .5f * (g.tab[a][i] + g.tab[a][i + 1]);
        }
        nt.back() = g.tab[a].back();
        g.tab[a].swap(nt);
        g.k[a]++;
    }
    // closed form available?  every boundary must equal wmin + i*w exactly (two separately rounded operations, as the kernels compute it)
    g.uniform = 1;
    for (int a = 0; a < 3; a++)
    {
        size_t n = g.tab[a].size() - 1;
        g.w[a] = g.tab[a][1] - g.tab[a][0];
        for (size_t i = 0; i <= n; i++)
        {
            volatile float prod = (float)i * g.w[a];
            volatile float v = prod + g.wmin[a];
            if (v != g.tab[a][i]) { g.uniform = 0; break; }
        }
    }
    return g.D >= 1 && g.D <= 32 && g.k[0] <= 11 && g.k[1] <= 11 && g.k[2] <= 11;
}
} // namespace bm
