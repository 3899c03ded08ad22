// bm_math.cuh -- arithmetic shared by the build and trace kernels (host+device inline so tests/hostcheck can run the very
// same source on the CPU against the oracle; the product only ever calls these from kernels).
//
// Bit-exactness rules (SURVEY.md section 7 "Hard parts"): every expression whose rounding can change a key, a traversal
// decision or a reported t is written in the same operation order as the reference and compiled WITHOUT fused
// multiply-add (-fmad=false for nvcc, -ffp-contract=off for g++); division and sqrt are IEEE (no --use_fast_math).
#pragma once
#include <stdint.h>
#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define BM_HD __host__ __device__ __forceinline__
#else
#define BM_HD inline
#endif

namespace bm
{
typedef uint32_t u32;
typedef int32_t i32;
typedef uint64_t u64;

struct V3 { float x, y, z; };
BM_HD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
BM_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
BM_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
BM_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
BM_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
BM_HD float get(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// glm::dot(vec3) = (x*x' + y*y') + z*z'   (glm/detail/func_geometric.inl:54-61)
BM_HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// glm::cross (func_geometric.inl:74-85)
BM_HD V3 cross3(V3 x, V3 y) { return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
// CudaComon.cuh:71-76 `a < b ? a : b` -- not fminf: a NaN operand must behave as in the reference
BM_HD float fmin_(float a, float b) { return a < b ? a : b; }
BM_HD float fmax_(float a, float b) { return a > b ? a : b; }

// CudaComon.cuh:158-172 bmBoxRayIntersect: entry distance clamped to >= 0, FLT_MAX on a miss or a box behind the origin
BM_HD float box_ray(V3 bMin, V3 bMax, V3 orig, V3 invDir)
{
    V3 tMin = (bMin - orig) * invDir;
    V3 tMax = (bMax - orig) * invDir;
    float ftmax = fmin_(fmax_(tMin.x, tMax.x), fmin_(fmax_(tMin.y, tMax.y), fmax_(tMin.z, tMax.z)));
    if (ftmax < 0.f) return FLT_MAX;
    float ftmin = fmax_(fmin_(tMin.x, tMax.x), fmax_(fmin_(tMin.y, tMax.y), fmin_(tMin.z, tMax.z)));
    float dist = fmax_(0.f, ftmin);
    return ftmax >= ftmin ? dist : FLT_MAX;
}

// Conservative cull used on the tight (triangle-union) boxes: does the infinite LINE through orig along dir meet the box?
// No sign test on purpose: the reference accepts hits with negative t (CudaComon.cuh:149-150, BuildTree.cu:419), so a leaf
// whose triangles lie behind the origin must not be culled in reference mode.  `grow` pads the box by a few ulps so that a
// hit found numerically by the triangle test can never be lost to rounding in this test.
BM_HD bool line_hits_box(V3 bMin, V3 bMax, V3 orig, V3 invDir)
{
    V3 tMin = (bMin - orig) * invDir;
    V3 tMax = (bMax - orig) * invDir;
    float ftmax = fminf(fmaxf(tMin.x, tMax.x), fminf(fmaxf(tMin.y, tMax.y), fmaxf(tMin.z, tMax.z)));
    float ftmin = fmaxf(fminf(tMin.x, tMax.x), fmaxf(fminf(tMin.y, tMax.y), fminf(tMin.z, tMax.z)));
    // fminf/fmaxf drop NaNs (0*inf when the origin lies on a slab plane of a direction-parallel box): conservative
    return !(ftmax < ftmin);
}

// CudaComon.cuh:117-155 bmTriIntersect with the two edges precomputed (e1 = v1 - v0, e2 = v2 - v0 are the same float
// subtractions the reference performs first, so every later bit is identical).
BM_HD float tri_intersect(V3 orig, V3 dir, V3 v0, V3 e1, V3 e2, float& u, float& v)
{
    V3 pvec = cross3(dir, e2);
    float det = dot3(e1, pvec);
    float invDet = 1.f / det;
    V3 tvec = orig - v0;
    u = dot3(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return FLT_MAX;
    V3 qvec = cross3(tvec, e1);
    v = dot3(dir, qvec) * invDet;
    if (v < 0 || v + u > 1) return FLT_MAX;
    return dot3(e2, qvec) * invDet;
}

BM_HD float fma_(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// Conservative pre-test in front of tri_intersect: true = tri_intersect WOULD return FLT_MAX ("no hit"), decided without its division
// and in fused arithmetic.  It may answer false for a triangle the exact test then rejects; it never answers true for one the exact test
// accepts, so results do not depend on it.
// tri_intersect accepts when u = fl(U * fl(1 / D)) lies in [0, 1], v = fl(V * fl(1 / D)) >= 0 and fl(u + v) <= 1 with
//   D = e1 . (dir x e2),  U = tvec . (dir x e2),  V = dir . (tvec x e1),
// all evaluated in unfused float arithmetic; each product / quotient contributes a relative error of at most 2^-24, hence acceptance
// implies  0 <= U / D <= 1 + 3 eps,  0 <= V / D  and  (U + V) / D <= 1 + 5 eps  for the UNFUSED values.  The fused values computed here
// differ from those by at most 9 eps * B with B the sum of the absolute values of all products involved (B = sum_i |x_i| (|a_j b_k| +
// |a_k b_j|) for a triple product): 2 eps from each cross-product component in either arithmetic, 3 eps from the dot product.  With
// s = sign(D') the tests below therefore use a margin of K (B_U + B_D) etc. with K = 4e-6 = 67 eps, seven times the bound; they also hold
// when D' and D disagree in sign (then both are below the bound in magnitude, and so must U and V be for the exact test to accept).
// A NaN anywhere answers false (every comparison fails), leaving the decision to the exact test.
BM_HD bool tri_reject(V3 orig, V3 dir, V3 v0, V3 e1, V3 e2)
{
    const float K = 4e-6f;
    // dir x e2 and its termwise absolute bound
    const float px = fma_(dir.y, e2.z, -(e2.y * dir.z)), py = fma_(dir.z, e2.x, -(e2.z * dir.x)), pz = fma_(dir.x, e2.y, -(e2.x * dir.y));
    const float ax = fma_(fabsf(dir.y), fabsf(e2.z), fabsf(e2.y * dir.z)), ay = fma_(fabsf(dir.z), fabsf(e2.x), fabsf(e2.z * dir.x)),
                az = fma_(fabsf(dir.x), fabsf(e2.y), fabsf(e2.x * dir.y));
    const float D = fma_(e1.z, pz, fma_(e1.y, py, e1.x * px));
    const float bD = fma_(fabsf(e1.z), az, fma_(fabsf(e1.y), ay, fabsf(e1.x) * ax));
    const V3 t = orig - v0;
    const float U = fma_(t.z, pz, fma_(t.y, py, t.x * px));
    const float bU = fma_(fabsf(t.z), az, fma_(fabsf(t.y), ay, fabsf(t.x) * ax));
    const bool neg = D < 0.f;
    const float aD = fabsf(D), a = neg ? -U : U, mU = K * (bU + bD);
    if (a < -mU || a > aD + mU) return true;
    // tvec x e1
    const float qx = fma_(t.y, e1.z, -(e1.y * t.z)), qy = fma_(t.z, e1.x, -(e1.z * t.x)), qz = fma_(t.x, e1.y, -(e1.x * t.y));
    const float cx = fma_(fabsf(t.y), fabsf(e1.z), fabsf(e1.y * t.z)), cy = fma_(fabsf(t.z), fabsf(e1.x), fabsf(e1.z * t.x)),
                cz = fma_(fabsf(t.x), fabsf(e1.y), fabsf(e1.x * t.y));
    const float V = fma_(dir.z, qz, fma_(dir.y, qy, dir.x * qx));
    const float bV = fma_(fabsf(dir.z), cz, fma_(fabsf(dir.y), cy, fabsf(dir.x) * cx));
    const float b = neg ? -V : V;
    if (b < -K * (bV + bD)) return true;
    return a + b > aD + K * (bU + bV + bD);
}

// Hash.cu:17-33 bmHash (Fletcher-16 over the 4 little-endian bytes), :41-47 bmHash3(uint3), :56-59 bmMap
BM_HD u32 hash32(u32 h)
{
    u32 sum1 = 0, sum2 = 0;
    for (int i = 0; i < 4; i++) { sum1 = (sum1 + ((h >> (8 * i)) & 0xFFu)) % 255u; sum2 = (sum2 + sum1) % 255u; }
    return (sum2 << 8) | sum1;
}
BM_HD u32 hash3(u32 x, u32 y, u32 z) { return (hash32(x) + hash32(y) + hash32(z)) % 65536u; }
BM_HD i32 map_cell(float f) { return (i32)floorf(f * (1.f / 0.03f)); }

// ---------------------------------------------------------------------------------------------------------------------
// The implicit grid behind the reference's kd-tree.  With world box W and midpoint splits on
// axes x,y,z,x,... the node at depth d is one cell of a 2^nx(d) x 2^ny(d) x 2^nz(d) grid; leaves all sit at depth D.
// tab[a][i], i = 0..2^k[a], are the cell boundaries of axis a at leaf resolution, produced on the host by the same
// recursive float midpoint the reference uses (BuildTree.cu:216), so every ancestor boundary is tab[a][i << shift].
// ---------------------------------------------------------------------------------------------------------------------
struct Grid
{
    const float* tab[3];
    int k[3];       // bits per axis at leaf depth
    int D;          // leaf depth = k[0]+k[1]+k[2] <= 32
    // When every table entry equals wmin + i*w bit for bit (true for the reference's +-30 world: all boundaries are exact floats),
    // `uniform` is set and boundaries are computed instead of loaded (checked entry by entry on the host, bm_grid_host.h).
    int uniform;
    float wmin[3], w[3];
    float wext[3];  // world extent per axis (wmax - wmin)
    // bits of the in-cell Morton sub-key appended to the leaf code in the sort key (3, 6, ... 15; chosen per build, tree.cu)
    int subBits;
};

BM_HD int axis_bits(int depth, int axis) { return (depth + 2 - axis) / 3; }   // splits on `axis` among depths 0..depth-1

BM_HD u32 spread3(u32 v)   // 10 bits -> every third bit
{
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
BM_HD u32 compact3(u32 v)
{
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030C30C3u;
    v = (v | (v >> 4)) & 0x0300F00Fu;
    v = (v | (v >> 8)) & 0x030000FFu;
    v = (v | (v >> 16)) & 0x3FFu;
    return v;
}

// Path code of the depth-`depth` node with per-axis cell indices (ix,iy,iz) (each with axis_bits(depth,a) bits): the bits
// of the root-to-node path, first split (x) at the most significant position.  Works for depth <= 32.
BM_HD u32 encode_path(int depth, u32 ix, u32 iy, u32 iz)
{
    int r = depth % 3;
    if (r == 0) return (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
    if (r == 1) { u32 c = (spread3(ix >> 1) << 2) | (spread3(iy) << 1) | spread3(iz); return (c << 1) | (ix & 1u); }
    u32 c = (spread3(ix >> 1) << 2) | (spread3(iy >> 1) << 1) | spread3(iz);
    return (c << 2) | ((ix & 1u) << 1) | (iy & 1u);
}
BM_HD void decode_path(int depth, u32 p, u32& ix, u32& iy, u32& iz)
{
    int r = depth % 3;
    if (r == 0) { ix = compact3(p >> 2); iy = compact3(p >> 1); iz = compact3(p); return; }
    if (r == 1) { u32 c = p >> 1; ix = (compact3(c >> 2) << 1) | (p & 1u); iy = compact3(c >> 1); iz = compact3(c); return; }
    u32 c = p >> 2;
    ix = (compact3(c >> 2) << 1) | ((p >> 1) & 1u); iy = (compact3(c >> 1) << 1) | (p & 1u); iz = compact3(c);
}

#if defined(__CUDA_ARCH__)
#define BM_LDG(p) __ldg(p)
#else
#define BM_LDG(p) (*(p))
#endif

// boundary i (leaf resolution) of axis a
BM_HD float grid_at(const Grid& g, int a, u32 i) { return g.uniform ? (float)i * g.w[a] + g.wmin[a] : BM_LDG(g.tab[a] + i); }

// Box of the depth-`depth` cell (ix,iy,iz)
BM_HD void cell_box(const Grid& g, int depth, u32 ix, u32 iy, u32 iz, V3& mn, V3& mx)
{
    int sx = g.k[0] - axis_bits(depth, 0), sy = g.k[1] - axis_bits(depth, 1), sz = g.k[2] - axis_bits(depth, 2);
    mn.x = grid_at(g, 0, ix << sx); mx.x = grid_at(g, 0, (ix + 1) << sx);
    mn.y = grid_at(g, 1, iy << sy); mx.y = grid_at(g, 1, (iy + 1) << sy);
    mn.z = grid_at(g, 2, iz << sz); mx.z = grid_at(g, 2, (iz + 1) << sz);
}
} // namespace bm
