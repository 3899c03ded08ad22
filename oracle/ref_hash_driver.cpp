// TEST INFRASTRUCTURE ONLY (see oracle/README.md). Not part of the product.
//
// C hooks onto the reference's spatial-hash key functions (Raytracer/Hash.cu:17-64), compiled from the
// reference source where it lies (CUDA 0 mode) into oracle/_ref/libbeam_ref_hash.so.  A separate library because
// Hash.cu and BuildTree.cu both include the guard-less BoxTriangle.cuh and would collide at link time.
#include "CudaComon.cuh"

u32 bmHash(u32 h);
u32 bmHash3(uint3 v);
u32 bmHash3(const vec3& v);
i32 bmMap(float f);

extern "C"
{
    u32 bmGetCellSize(); u32 bmGetMaxHashElements();
    u32 ref_hash(u32 h) { return bmHash(h); }
    u32 ref_hash3(u32 x, u32 y, u32 z) { return bmHash3(make_uint3(x, y, z)); }
    u32 ref_hash3f(float x, float y, float z) { return bmHash3(vec3(x, y, z)); }
    i32 ref_map(float f) { return bmMap(f); }
    u32 ref_cell_size() { return bmGetCellSize(); }
    u32 ref_max_hash_elements() { return bmGetMaxHashElements(); }
}
