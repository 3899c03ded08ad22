// bm_wide.cuh -- the compressed 4-wide node: 64 bytes, 64-byte aligned, one record per internal node of the binary radix tree.
//
// The record of binary node b describes TWO levels of the tree below b: its children A (lower half) and B (upper half) are absorbed when
// they are internal, so a visit sees up to four grandchildren ("slots" 0,1 below A and 2,3 below B; a child that is a leaf occupies the
// first slot of its pair and leaves the second empty).  A traversal that enters at the root and always continues at a slot's reference
// touches every second level only: half the dependent memory round trips of the binary walk, 64 instead of 96 bytes per visit.
//
// What the reference's traversal needs from an internal node is its implicit cell box, split axis and split plane (BuildTree.cu:436-481):
// all three are a pure function of the node's position in the kd-tree, so the record stores only that position -- the integer cell
// coordinates at the node's depth -- for b, A and B, and the kernels rebuild the reference's values exactly (cell_box / grid_at: the same
// float operations that filled the grid tables).  The only geometry stored is the culling box of each slot (union of the WHOLE triangles
// below it, as before), quantised to 8 bits per coordinate on a per-node grid and rounded OUTWARDS: a culling box never decides a
// result, it only prunes, so precision costs visits, not correctness.
//
//   q0 = ( cell(b), cell(A), cell(B), len(b) | len(A) << 8 | len(B) << 16 | expX << 24 )      len = 0xFF: child not absorbed
//        cell = ix | iy << k0 | iz << (k0 + k1): cell coordinates at the node's own depth (k = leaf-level bits per axis, Grid::k)
//   q1 = ( ref0, ref1, ref2, ref3 )            bit 31 = leaf index, 0x7FFFFFFF = empty slot, else index of the slot's own record
//   q2 = ( LX, LY, LZ, HX )                    byte s of LX = quantised lo.x of slot s, ...
//   q3 = ( HY, HZ, orgX | orgY << 16, orgZ | expY << 16 | expZ << 24 )
//
// Quantisation grid of axis a:  origin_a = (cellMin_a - 3 ext_a) + org_a * (ext_a * 2^-14)   (cell of b, ext = cellMax - cellMin)
//                               step_a   = 2^(exp_a - 64)
//                               value(q) = fma(2^23 + q, step_a, origin_a - 2^23 step_a)      one PRMT + one FFMA per coordinate
// exp bit 7 set = the grid of this axis is anchored at the world box instead of the cell (a slot box reaching more than three cells below the
// node's cell: triangles much larger than the cell); exp = 0xFF = unbounded on this axis (every slot spans (-1e30, 1e30) there: boxes
// reaching more than a world extent outside the world).  wide_encode verifies with the decoder's own arithmetic that every decoded bound
// lies on the outer side of the exact one.
#pragma once
#include <float.h>
#include <math.h>
#include "bm_math.cuh"
#include "../../include/beam_b200.h"

namespace bm
{
constexpr u32 WREF_EMPTY = 0x7FFFFFFFu;
constexpr u32 WLEN_NONE = 0xFFu;
constexpr int WNODE_U4 = 4;                 // uint4 per record

// cell coordinates <-> packed word (field widths = leaf-level bits per axis)
BM_HD u32 wide_pack_cell(const Grid& g, u32 ix, u32 iy, u32 iz) { return ix | (iy << g.k[0]) | (iz << (g.k[0] + g.k[1])); }
BM_HD void wide_unpack_cell(const Grid& g, u32 p, u32& ix, u32& iy, u32& iz)
{
    ix = p & ((1u << g.k[0]) - 1u); iy = (p >> g.k[0]) & ((1u << g.k[1]) - 1u); iz = p >> (g.k[0] + g.k[1]);
}

// the reference node at depth `len` with cell coordinates `packed`: its box, split axis and split plane (BuildTree.cu:436-443)
struct WCell { V3 mn, mx; float split; int axis; };
BM_HD WCell wide_cell(const Grid& g, u32 packed, int len)
{
    WCell c;
    u32 ix, iy, iz; wide_unpack_cell(g, packed, ix, iy, iz);
    // cell_box with the three axis_bits() folded into one division: splits on x, y, z among depths 0..len-1 = q + (r > 0), q + (r > 1), q
    const int q = len / 3, r = len - 3 * q;
    const int sx = g.k[0] - q - (r > 0 ? 1 : 0), sy = g.k[1] - q - (r > 1 ? 1 : 0), sz = g.k[2] - q;
    c.mn.x = grid_at(g, 0, ix << sx); c.mx.x = grid_at(g, 0, (ix + 1) << sx);
    c.mn.y = grid_at(g, 1, iy << sy); c.mx.y = grid_at(g, 1, (iy + 1) << sy);
    c.mn.z = grid_at(g, 2, iz << sz); c.mx.z = grid_at(g, 2, (iz + 1) << sz);
    c.axis = r;
    // the reference's own expression (BuildTree.cu:443); equal to the grid boundary between the two halves, which the host tables
    // produce with this very midpoint (bm_grid_host.h)
    c.split = .5f * (get(c.mx, r) + get(c.mn, r));
    return c;
}

// decoder state of one record: per axis the step and the two addends (lo / hi differ only on "unbounded" axes)
struct WFrame { float step[3], addLo[3], addHi[3]; };

BM_HD float wide_pow2(int e) { union { u32 u; float f; } c; c.u = (u32)(e + 127) << 23; return c.f; }      // 2^e, e in [-126, 127]

// anchor of axis a's quantisation grid: (base, unit) with origin = base + org * unit
//   normal       : base = cellMin - 3 ext, unit = ext * 2^-14            (reaches three cells below the node's cell, resolution ext / 16384)
//   world (bit 7): base = worldMin - worldExt, unit = worldExt * 2^-15   (boxes of triangles much larger than the cell; reaches one world
//                                                                         extent outside the world box)
BM_HD void wide_anchor(const Grid& g, int a, float cellMn, float cellMx, bool world, float& base, float& unit)
{
    if (world)
    {
        float wext = g.wext[a];
        base = g.wmin[a] - wext; unit = wext * 3.0517578125e-05f;
    }
    else
    {
        float ext = cellMx - cellMn;
        base = cellMn - 3.f * ext; unit = ext * 6.103515625e-05f;
    }
}

BM_HD WFrame wide_frame(const Grid& g, const V3& cellMn, const V3& cellMx, u32 orgXY, u32 orgZexp, u32 expX)
{
    WFrame f;
    const u32 org[3] = { orgXY & 0xFFFFu, orgXY >> 16, orgZexp & 0xFFFFu };
    const u32 ex[3] = { expX & 0xFFu, (orgZexp >> 16) & 0xFFu, orgZexp >> 24 };
    const float mn[3] = { cellMn.x, cellMn.y, cellMn.z }, mx[3] = { cellMx.x, cellMx.y, cellMx.z };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int a = 0; a < 3; a++)
    {
        if (ex[a] == 0xFFu) { f.step[a] = 0.f; f.addLo[a] = -1e30f; f.addHi[a] = 1e30f; continue; }
        float base, unit;
        wide_anchor(g, a, mn[a], mx[a], (ex[a] & 0x80u) != 0u, base, unit);
        float origin = (float)org[a] * unit + base;
        float step = wide_pow2((int)(ex[a] & 0x7Fu) - 64);
        f.step[a] = step;
        f.addLo[a] = f.addHi[a] = origin - 8388608.f * step;
    }
    return f;
}

// value of quantised coordinate q (0..255): host form of the kernels' PRMT + FFMA
BM_HD float wide_value(u32 q, float step, float add)
{
    union { u32 u; float f; } c; c.u = 0x4B000000u | q;         // 2^23 + q
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(c.f, step, add);
#else
    return fmaf(c.f, step, add);
#endif
}

// byte s of `word` as the float 2^23 + byte (one PRMT on the device)
BM_HD float wide_magic(u32 word, int s)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u + (u32)s));
#else
    union { u32 u; float f; } c; c.u = 0x4B000000u | ((word >> (8 * s)) & 0xFFu); return c.f;
#endif
}
BM_HD float wide_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

struct WQ { u32 x, y, z, w; };                    // one quarter of a record (uint4 on the device)

// ---- encoder (build side) -------------------------------------------------------------------------------------------------------
struct WSlotBox { float lo[3], hi[3]; };

// Writes the 64-byte record.  slots with ref == WREF_EMPTY are ignored.  Boxes are the exact (padded) unions the hierarchy pass carries.
BM_HD void wide_encode(const Grid& g, u32 cell0, u32 len0, u32 cellA, u32 lenA, u32 cellB, u32 lenB, const u32 ref[4], const WSlotBox box[4],
                       WQ* out)
{
    WCell c0 = wide_cell(g, cell0, (int)len0);
    const float cmn[3] = { c0.mn.x, c0.mn.y, c0.mn.z }, cmx[3] = { c0.mx.x, c0.mx.y, c0.mx.z };
    u32 org[3], ex[3], L[3] = { 0, 0, 0 }, H[3] = { 0, 0, 0 };
    for (int a = 0; a < 3; a++)
    {
        float flo = FLT_MAX, fhi = -FLT_MAX;
        for (int s = 0; s < 4; s++) if (ref[s] != WREF_EMPTY) { flo = fminf(flo, box[s].lo[a]); fhi = fmaxf(fhi, box[s].hi[a]); }
        bool ok = false;
        int e = 0; u32 o = 0, worldBit = 0; float step = 0.f, add = 0.f;
        for (int attempt = 0; attempt < 2 && !ok; attempt++)
        {
            float base, unit;
            wide_anchor(g, a, cmn[a], cmx[a], attempt == 1, base, unit);
            if (!(flo >= base && fhi >= flo && fhi < 1e29f)) continue;
            // smallest power-of-two step that spans the node's box in ~248 steps, then widened until the decoder itself confirms coverage
            int fe; frexpf(fmaxf(fhi - flo, 1e-30f) * (1.f / 248.f), &fe);
            e = fe < -60 ? -60 : fe;
            for (int tries = 0; tries < 48 && e <= 20; tries++, e++)
            {
                step = wide_pow2(e);
                float target = flo - 2.f * step;
                if (target < base) target = base;
                float fo = floorf((target - base) / unit);
                o = fo < 0.f ? 0u : (fo > 65535.f ? 65535u : (u32)fo);
                for (int k = 0; k < 4; k++)
                {
                    float origin = (float)o * unit + base;
                    add = origin - 8388608.f * step;
                    if (wide_value(0, step, add) <= flo || o == 0) break;
                    o--;
                }
                if (wide_value(0, step, add) <= flo && wide_value(255, step, add) >= fhi) { ok = true; worldBit = attempt ? 0x80u : 0u; break; }
            }
        }
        if (!ok)
        {
            org[a] = 0; ex[a] = 0xFFu;
            for (int s = 0; s < 4; s++) { H[a] |= 255u << (8 * s); }
            continue;
        }
        org[a] = o; ex[a] = (u32)(e + 64) | worldBit;
        const float dec0 = wide_value(0, step, add);
        for (int s = 0; s < 4; s++)
        {
            u32 ql = 255u, qh = 0u;                                       // empty slot: inverted interval
            if (ref[s] != WREF_EMPTY)
            {
                float fl = floorf((box[s].lo[a] - dec0) / step), fh = ceilf((box[s].hi[a] - dec0) / step);
                ql = fl < 0.f ? 0u : (fl > 255.f ? 255u : (u32)fl);
                qh = fh < 0.f ? 0u : (fh > 255.f ? 255u : (u32)fh);
                while (ql > 0u && wide_value(ql, step, add) > box[s].lo[a]) ql--;
                while (qh < 255u && wide_value(qh, step, add) < box[s].hi[a]) qh++;
            }
            L[a] |= ql << (8 * s); H[a] |= qh << (8 * s);
        }
    }
    out[0].x = cell0; out[0].y = cellA; out[0].z = cellB; out[0].w = len0 | (lenA << 8) | (lenB << 16) | (ex[0] << 24);
    out[1].x = ref[0]; out[1].y = ref[1]; out[1].z = ref[2]; out[1].w = ref[3];
    out[2].x = L[0]; out[2].y = L[1]; out[2].z = L[2]; out[2].w = H[0];
    out[3].x = H[1]; out[3].y = H[2]; out[3].z = org[0] | (org[1] << 16); out[3].w = org[2] | (ex[1] << 16) | (ex[2] << 24);
}

// ---- one traversal step (shared by the trace kernels and tests/hostcheck) ---------------------------------------------------------
// Culling per slot, on the decoded (outward-rounded) box of the whole triangles below it:
//   both modes : the ray's LINE must meet the box (a triangle can only be hit where its box is);
//   REFERENCE  : the box grown by one leaf cell must not lie entirely behind the origin.  A leaf is tested only if the forward ray enters
//                its CELL (BuildTree.cu:398-406), every leaf cell of the subtree overlaps one of its triangles, hence lies inside the grown
//                box; fm = leafExt * |1/d| is how far that growth moves the exit distances.  (The triangles themselves may be hit at
//                t < 0 -- the reference accepts that -- so the un-grown box must NOT be tested for direction.)
//   CLOSEST    : the box must overlap [0, best t].
// Order (REFERENCE): the reference orders the two children of a kd node by the side of the split plane on which the ray ENTERS the node's
// cell (BuildTree.cu:440-481).  The record holds three such nodes -- b and its absorbed children A, B -- and each decision is evaluated
// with the reference's own arithmetic on the rebuilt cell, but only where it can matter: when slots on both sides of it survive.  A cell
// the ray misses drops its slots, as the reference's `boxDist == FLT_MAX` does.
// Returns the number of references written to push[] in PUSH order: the entry to visit next is the last one.
// ic = 1/d clamped to +-1e30 (never inf or NaN), oi = o * ic; inv = the unclamped reciprocal the reference's slab test uses.
template <int MODE>
BM_HD int wide_step(const Grid& g, WQ q0, WQ q1, WQ q2, WQ q3, V3 o, V3 d, V3 inv, V3 ic, V3 oi, V3 fm, float tbest, u32 push[4])
{
    const WCell c0 = wide_cell(g, q0.x, (int)(q0.w & 0xFFu));
    float boxDist0 = 0.f;
    if (MODE == BM_TRACE_REFERENCE)
    {
        boxDist0 = box_ray(c0.mn, c0.mx, o, inv);                                                  // the node's cell, :398
        if (boxDist0 == FLT_MAX) return 0;
    }
    const WFrame f = wide_frame(g, c0.mn, c0.mx, q3.z, q3.w, q0.w >> 24);
    const u32 refs[4] = { q1.x, q1.y, q1.z, q1.w };
    bool h[4]; float tn[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < 4; s++)
    {
        const float lx = wide_fma(wide_magic(q2.x, s), f.step[0], f.addLo[0]), hx = wide_fma(wide_magic(q2.w, s), f.step[0], f.addHi[0]);
        const float ly = wide_fma(wide_magic(q2.y, s), f.step[1], f.addLo[1]), hy = wide_fma(wide_magic(q3.x, s), f.step[1], f.addHi[1]);
        const float lz = wide_fma(wide_magic(q2.z, s), f.step[2], f.addLo[2]), hz = wide_fma(wide_magic(q3.y, s), f.step[2], f.addHi[2]);
        const float x0 = wide_fma(lx, ic.x, -oi.x), x1 = wide_fma(hx, ic.x, -oi.x);
        const float y0 = wide_fma(ly, ic.y, -oi.y), y1 = wide_fma(hy, ic.y, -oi.y);
        const float z0 = wide_fma(lz, ic.z, -oi.z), z1 = wide_fma(hz, ic.z, -oi.z);
        const float fx = fmaxf(x0, x1), fy = fmaxf(y0, y1), fz = fmaxf(z0, z1);
        const float tnear = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
        const float tfar = fminf(fminf(fx, fy), fz);
        bool hit = refs[s] != WREF_EMPTY && !(tfar < tnear);
        if (MODE == BM_TRACE_REFERENCE) hit = hit && !(fminf(fminf(fx + fm.x, fy + fm.y), fz + fm.z) < 0.f);
        else hit = hit && !(tfar < 0.f) && !(tnear > tbest);
        h[s] = hit; tn[s] = tnear;
    }
    int n = 0;
    if (MODE == BM_TRACE_REFERENCE)
    {
        bool a01 = true, b23 = true, aFirst = true;          // lower slot / lower half first
        // the cell of an absorbed child is rebuilt only where both of its slots survive; a lane rarely needs both children, so the first
        // pass serves whichever one it needs (one instruction stream for the warp instead of two half-empty ones) and the second the rest
        const bool needA = h[0] && h[1], needB = h[2] && h[3];
        if (needA || needB)
        {
            const WCell cX = wide_cell(g, needA ? q0.y : q0.z, (int)((q0.w >> (needA ? 8 : 16)) & 0xFFu));
            const float bd = box_ray(cX.mn, cX.mx, o, inv);
            const bool lowFirst = get(o, cX.axis) + bd * get(d, cX.axis) < cX.split;                // :443-444
            if (needA) { if (bd == FLT_MAX) h[0] = h[1] = false; else a01 = lowFirst; }
            else { if (bd == FLT_MAX) h[2] = h[3] = false; else b23 = lowFirst; }
        }
        if (needA && needB)
        {
            const WCell cB = wide_cell(g, q0.z, (int)((q0.w >> 16) & 0xFFu));
            const float bd = box_ray(cB.mn, cB.mx, o, inv);
            if (bd == FLT_MAX) h[2] = h[3] = false;
            else b23 = get(o, cB.axis) + bd * get(d, cB.axis) < cB.split;
        }
        if ((h[0] || h[1]) && (h[2] || h[3])) aFirst = get(o, c0.axis) + boxDist0 * get(d, c0.axis) < c0.split;
        const u32 An = a01 ? refs[0] : refs[1], Af = a01 ? refs[1] : refs[0], Bn = b23 ? refs[2] : refs[3], Bf = b23 ? refs[3] : refs[2];
        const bool hAn = a01 ? h[0] : h[1], hAf = a01 ? h[1] : h[0], hBn = b23 ? h[2] : h[3], hBf = b23 ? h[3] : h[2];
        // far group first, inside a group the far slot first: the near slot of the near group is visited next
        const u32 r3 = aFirst ? Bf : Af, r2 = aFirst ? Bn : An, r1 = aFirst ? Af : Bf, r0 = aFirst ? An : Bn;
        const bool p3 = aFirst ? hBf : hAf, p2 = aFirst ? hBn : hAn, p1 = aFirst ? hAf : hBf, p0 = aFirst ? hAn : hBn;
        if (p3) push[n++] = r3;
        if (p2) push[n++] = r2;
        if (p1) push[n++] = r1;
        if (p0) push[n++] = r0;
    }
    else
    {
        // nearest box visited next: the (at most four) survivors sorted by entry distance, farthest first
        u32 r[4] = { refs[0], refs[1], refs[2], refs[3] };
        float k[4];
        for (int s = 0; s < 4; s++) k[s] = h[s] ? tn[s] : FLT_MAX;
#define BM_CSWAP(i, j) { if (k[i] < k[j]) { float tk = k[i]; k[i] = k[j]; k[j] = tk; u32 tr = r[i]; r[i] = r[j]; r[j] = tr; } }
        BM_CSWAP(0, 1) BM_CSWAP(2, 3) BM_CSWAP(0, 2) BM_CSWAP(1, 3) BM_CSWAP(1, 2)          // descending entry distance
#undef BM_CSWAP
        for (int s = 0; s < 4; s++) if (k[s] != FLT_MAX) push[n++] = r[s];
    }
    return n;
}
} // namespace bm
