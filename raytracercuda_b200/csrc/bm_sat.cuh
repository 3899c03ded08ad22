// bm_sat.cuh -- triangle / box separating-axis test with the reference's rounding (BoxTriangle.cuh:57-79,134-222) and
// the key-generation descent built on it.
//
// The boolean result is the AND of 13 axis tests, so the order in which they are evaluated is free; each individual
// test is computed with exactly the reference's expressions.
#pragma once
#include "bm_math.cuh"

namespace bm
{
struct TriSat { V3 v[3]; };
// The reference forms the edges from the box-relative vertices, e0 = (v1-c)-(v0-c) (BoxTriangle.cuh:155-163), so through
// rounding they depend on the box centre; sat_box() therefore recomputes them per box and TriSat carries only vertices.
BM_HD void trisat_init(TriSat& t, V3 a, V3 b, V3 c) { t.v[0] = a; t.v[1] = b; t.v[2] = c; }

#define BM_AX(pa, pb, rad) { float mn_, mx_; if ((pa) < (pb)) { mn_ = (pa); mx_ = (pb); } else { mn_ = (pb); mx_ = (pa); } if (mn_ > (rad) || mx_ < -(rad)) return 0; }

// triBoxOverlap(boxcenter, boxhalfsize, triverts), BoxTriangle.cuh:134-222. Returns 0/1.
BM_HD int sat_box(const TriSat& t, const float bc[3], const float hs[3])
{
    float v0[3] = { t.v[0].x - bc[0], t.v[0].y - bc[1], t.v[0].z - bc[2] };
    float v1[3] = { t.v[1].x - bc[0], t.v[1].y - bc[1], t.v[1].z - bc[2] };
    float v2[3] = { t.v[2].x - bc[0], t.v[2].y - bc[1], t.v[2].z - bc[2] };
    // box axes first (cheapest reject; FINDMINMAX :49-54, :195-208)
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        float mn = v0[k], mx = v0[k];
        if (v1[k] < mn) mn = v1[k];
        if (v1[k] > mx) mx = v1[k];
        if (v2[k] < mn) mn = v2[k];
        if (v2[k] > mx) mx = v2[k];
        if (mn > hs[k] || mx < -hs[k]) return 0;
    }
    float e0[3] = { v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2] };
    float e1[3] = { v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2] };
    float e2[3] = { v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2] };
    float fex, fey, fez, p0, p1, p2, rad;
    const int X = 0, Y = 1, Z = 2;

    fex = fabsf(e0[X]); fey = fabsf(e0[Y]); fez = fabsf(e0[Z]);
    p0 = e0[Z] * v0[Y] - e0[Y] * v0[Z]; p2 = e0[Z] * v2[Y] - e0[Y] * v2[Z]; rad = fez * hs[Y] + fey * hs[Z]; BM_AX(p0, p2, rad)       // AXISTEST_X01
    p0 = -e0[Z] * v0[X] + e0[X] * v0[Z]; p2 = -e0[Z] * v2[X] + e0[X] * v2[Z]; rad = fez * hs[X] + fex * hs[Z]; BM_AX(p0, p2, rad)     // AXISTEST_Y02
    p1 = e0[Y] * v1[X] - e0[X] * v1[Y]; p2 = e0[Y] * v2[X] - e0[X] * v2[Y]; rad = fey * hs[X] + fex * hs[Y]; BM_AX(p2, p1, rad)       // AXISTEST_Z12

    fex = fabsf(e1[X]); fey = fabsf(e1[Y]); fez = fabsf(e1[Z]);
    p0 = e1[Z] * v0[Y] - e1[Y] * v0[Z]; p2 = e1[Z] * v2[Y] - e1[Y] * v2[Z]; rad = fez * hs[Y] + fey * hs[Z]; BM_AX(p0, p2, rad)       // AXISTEST_X01
    p0 = -e1[Z] * v0[X] + e1[X] * v0[Z]; p2 = -e1[Z] * v2[X] + e1[X] * v2[Z]; rad = fez * hs[X] + fex * hs[Z]; BM_AX(p0, p2, rad)     // AXISTEST_Y02
    p0 = e1[Y] * v0[X] - e1[X] * v0[Y]; p1 = e1[Y] * v1[X] - e1[X] * v1[Y]; rad = fey * hs[X] + fex * hs[Y]; BM_AX(p0, p1, rad)       // AXISTEST_Z0

    fex = fabsf(e2[X]); fey = fabsf(e2[Y]); fez = fabsf(e2[Z]);
    p0 = e2[Z] * v0[Y] - e2[Y] * v0[Z]; p1 = e2[Z] * v1[Y] - e2[Y] * v1[Z]; rad = fez * hs[Y] + fey * hs[Z]; BM_AX(p0, p1, rad)       // AXISTEST_X2
    p0 = -e2[Z] * v0[X] + e2[X] * v0[Z]; p1 = -e2[Z] * v1[X] + e2[X] * v1[Z]; rad = fez * hs[X] + fex * hs[Z]; BM_AX(p0, p1, rad)     // AXISTEST_Y1
    p1 = e2[Y] * v1[X] - e2[X] * v1[Y]; p2 = e2[Y] * v2[X] - e2[X] * v2[Y]; rad = fey * hs[X] + fex * hs[Y]; BM_AX(p2, p1, rad)       // AXISTEST_Z12

    // triangle plane, planeBoxOverlap :57-79 with normal = CROSS(e0, e1) :214
    float nrm[3] = { e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0] };
    float vmin[3], vmax[3];
#pragma unroll
    for (int q = 0; q < 3; q++)
    {
        float v = v0[q];
        if (nrm[q] > 0.0f) { vmin[q] = -hs[q] - v; vmax[q] = hs[q] - v; }
        else               { vmin[q] = hs[q] - v;  vmax[q] = -hs[q] - v; }
    }
    if (nrm[0] * vmin[0] + nrm[1] * vmin[1] + nrm[2] * vmin[2] > 0.0f) return 0;
    if (nrm[0] * vmax[0] + nrm[1] * vmax[1] + nrm[2] * vmax[2] >= 0.0f) return 1;
    return 0;
}

// Leaf-resolution index range [lo,hi] on one axis of the cells whose box-axis test (FINDMINMAX branch above) passes for a
// triangle spanning [vmin,vmax] on that axis.  fl(v - c) is monotone in v and in c, so the passing cells form one
// interval; it is found from a float estimate and corrected with the exact test expressions.  Cells outside it can
// never be in the reference's leaf set for this triangle, which makes it a safe pruning bound for the descent.
BM_HD void axis_range(const Grid& g, int ax, float vmin, float vmax, int& lo, int& hi)
{
    const int n = 1 << g.k[ax];
    float w0 = grid_at(g, ax, 0), w1 = grid_at(g, ax, n);
    float inv = (float)n / (w1 - w0);
    // lower end: smallest i with !(fl(vmin - c_i) > h_i)
    float f = (vmin - w0) * inv;
    int i = f < 0.f ? 0 : (f >= (float)n ? n - 1 : (int)f);      // also maps NaN to n-1/0 harmlessly; corrected below
    for (;;)
    {
        if (i > 0)
        {
            float a = grid_at(g, ax, i - 1), b = grid_at(g, ax, i);
            float c = (b + a) * .5f, h = (b - a) * .5f;
            if (!(vmin - c > h)) { --i; continue; }
        }
        break;
    }
    for (;;)
    {
        if (i < n)
        {
            float a = grid_at(g, ax, i), b = grid_at(g, ax, i + 1);
            float c = (b + a) * .5f, h = (b - a) * .5f;
            if (vmin - c > h) { ++i; continue; }
        }
        break;
    }
    lo = i;
    // upper end: largest i with !(fl(vmax - c_i) < -h_i)
    f = (vmax - w0) * inv;
    i = f < 0.f ? 0 : (f >= (float)n ? n - 1 : (int)f);
    for (;;)
    {
        if (i < n - 1)
        {
            float a = grid_at(g, ax, i + 1), b = grid_at(g, ax, i + 2);
            float c = (b + a) * .5f, h = (b - a) * .5f;
            if (!(vmax - c < -h)) { ++i; continue; }
        }
        break;
    }
    for (;;)
    {
        if (i >= 0)
        {
            float a = grid_at(g, ax, i), b = grid_at(g, ax, i + 1);
            float c = (b + a) * .5f, h = (b - a) * .5f;
            if (vmax - c < -h) { --i; continue; }
        }
        break;
    }
    hi = i;
}

// ---- flat candidate test (the fast path of key generation) --------------------------------------------------------------
// Is (triangle, leaf cell) a pair of the reference tree?  The reference reaches a leaf only if ITS test passes on every
// ancestor box of the path (BuildTree.cu:216-226), and in float arithmetic the tests on large ancestor boxes are not implied
// by the test on the leaf (their rounding error grows with the box).  pair_test therefore walks the path upwards from the leaf
// and evaluates the reference's test on each box, but stops as soon as some vertex of the triangle lies inside the current box
// with margin `m` on all three axes.  From there on every larger box contains that vertex with at least the same margin (going
// to the parent moves one face outwards, none inwards), and a vertex inside a box with margin m >= ~8 * 2^-24 * (box size +
// |v - c|) makes each of the 13 float tests pass with room to spare:
//   * box axes: the vertex itself is inside;
//   * edge-cross axes: for the projected vertex |p| <= fa*(hy-m) + fb*(hz-m) = rad - m*(fa+fb) with the SAME float fa, fb on both
//     sides, so only the ~5 roundings of p and rad (each <= 2^-24 relative) are in play; the third vertex projects onto one of the
//     two that the test uses up to 2^-23 * |edge|^2;
//   * plane: the triangle's plane passes through that vertex, so the box's extreme corners lie >= m*|n|_1 on either side.
// m = world extent * 2^-19 (1.1e-4 for the +-30 world, 2.6 x the requirement at the largest box).  Hence the result equals the
// reference's descent by construction; tests/test_gpu_build.py also checks it bit for bit against the descent kernel.
BM_HD bool pair_test(const Grid& g, const TriSat& t, u32 ix, u32 iy, u32 iz, float m)
{
    for (int d = g.D; d >= 1; --d)
    {
        int sx = g.k[0] - axis_bits(d, 0), sy = g.k[1] - axis_bits(d, 1), sz = g.k[2] - axis_bits(d, 2);
        V3 mn, mx; cell_box(g, d, ix >> sx, iy >> sy, iz >> sz, mn, mx);
        float bc[3] = { (mx.x + mn.x) * .5f, (mx.y + mn.y) * .5f, (mx.z + mn.z) * .5f };
        float hs[3] = { (mx.x - mn.x) * .5f, (mx.y - mn.y) * .5f, (mx.z - mn.z) * .5f };
        float hx = hs[0] - m, hy = hs[1] - m, hz = hs[2] - m;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (fabsf(t.v[k].x - bc[0]) <= hx && fabsf(t.v[k].y - bc[1]) <= hy && fabsf(t.v[k].z - bc[2]) <= hz) return true;
        if (!sat_box(t, bc, hs)) return false;
    }
    return true;
}

// Reference descent for ONE triangle (BuildTree.cu:191-255), restricted to cells inside the exact per-axis ranges:
// depth-first over the implicit kd nodes; a child is entered iff the reference's own test on the child box passes;
// reaching depth D emits the leaf path code.  `emit(code, ix, iy, iz)` is called once per (leaf, triangle) pair.  Because every
// level evaluates the reference's test on the reference's box, the emitted set equals the reference's leaf set by
// construction (no float-monotonicity assumption).  Returns the number of pairs.
template <class Emit>
BM_HD u32 descend_triangle(const Grid& g, V3 a, V3 b, V3 c, Emit emit)
{
    TriSat t; trisat_init(t, a, b, c);
    int lo[3], hi[3];
    axis_range(g, 0, fminf(a.x, fminf(b.x, c.x)), fmaxf(a.x, fmaxf(b.x, c.x)), lo[0], hi[0]);
    axis_range(g, 1, fminf(a.y, fminf(b.y, c.y)), fmaxf(a.y, fmaxf(b.y, c.y)), lo[1], hi[1]);
    axis_range(g, 2, fminf(a.z, fminf(b.z, c.z)), fmaxf(a.z, fmaxf(b.z, c.z)), lo[2], hi[2]);
    if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) return 0;

    // explicit stack of packed (depth:6 | ix:11 | iy:11 | iz:11); a pop replaces one node by at most two deeper ones,
    // so it never holds more than D+1 <= 33 entries
    u64 stack[36];
    int top = 0;
    stack[0] = 0;
    u32 count = 0;
    while (top >= 0)
    {
        u64 e = stack[top--];
        u32 d = (u32)(e >> 33), ix = (u32)(e >> 22) & 0x7FFu, iy = (u32)(e >> 11) & 0x7FFu, iz = (u32)e & 0x7FFu;
        if ((int)d == g.D) { emit(encode_path(g.D, ix, iy, iz), ix, iy, iz); ++count; continue; }
        int ax = d % 3;
        u32 nd = d + 1;
#pragma unroll
        for (int side = 0; side < 2; side++)
        {
            u32 cx = ix, cy = iy, cz = iz;
            if (ax == 0) cx = (ix << 1) | side; else if (ax == 1) cy = (iy << 1) | side; else cz = (iz << 1) | side;
            // prune by the leaf-resolution ranges (outcome-neutral, see axis_range)
            int shx = g.k[0] - axis_bits(nd, 0), shy = g.k[1] - axis_bits(nd, 1), shz = g.k[2] - axis_bits(nd, 2);
            if ((int)(((cx + 1) << shx) - 1) < lo[0] || (int)(cx << shx) > hi[0]) continue;
            if ((int)(((cy + 1) << shy) - 1) < lo[1] || (int)(cy << shy) > hi[1]) continue;
            if ((int)(((cz + 1) << shz) - 1) < lo[2] || (int)(cz << shz) > hi[2]) continue;
            V3 mn, mx; cell_box(g, nd, cx, cy, cz, mn, mx);
            // bc = (max+min)*.5f, hs = (max-min)*.5f   BuildTree.cu:221-225
            float bc[3] = { (mx.x + mn.x) * .5f, (mx.y + mn.y) * .5f, (mx.z + mn.z) * .5f };
            float hs[3] = { (mx.x - mn.x) * .5f, (mx.y - mn.y) * .5f, (mx.z - mn.z) * .5f };
            if (!sat_box(t, bc, hs)) continue;
            stack[++top] = ((u64)nd << 33) | ((u64)cx << 22) | ((u64)cy << 11) | (u64)cz;
        }
    }
    return count;
}
} // namespace bm
