#!/usr/bin/env bash
# TEST / BASELINE INFRASTRUCTURE ONLY.  Compiles the reference's host classes in their shipping CUDA-1 mode with nvcc for
# sm_100a, from the sources where they lie under $REF, and links them two ways (see oracle/ref_gpu_driver.cpp):
#   oracle/_ref/libbeam_refhost_b200.so  reference host code, unmodified, on top of this repo's libbeam_b200.so (drop-in proof)
#   oracle/_ref/libbeam_refcuda.so       reference host code + the reference's own kernels (BuildTree.cu, RTClear.cu) for B200
# The reference kernels need ONE portability fix to compile with CUDA 12.9 on LP64: atomicCAS has no
# `unsigned long` overload, so CudaComon.cuh:34 is patched (in the scratch copy only) to route through `unsigned long long`.
# They are built with --use_fast_math like the reference project (Raytracer.vcxproj FastMath=true).  RenderTarget.cpp
# (Windows/OpenGL) is replaced by oracle/ref_headless_rt.cpp.  Nothing is copied into the repository; outputs only in _ref/.
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
LIBB200="$HERE/../raytracercuda_b200/libbeam_b200.so"
if [ ! -d "$REF/Raytracer" ]; then echo "build_ref_cuda.sh: $REF/Raytracer not present; keeping prebuilt oracle/_ref if any"; exit 0; fi
if [ ! -f "$LIBB200" ]; then echo "build_ref_cuda.sh: build libbeam_b200.so first"; exit 1; fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
for f in "$REF"/Raytracer/*; do
  b="$(basename "$f")"
  [ "$b" = "CudaComon.cuh" ] && continue
  ln -s "$f" "$TMP/$b"
done
sed 's/return atomicCAS(lk, old, nw);/return (T)atomicCAS((unsigned long long*)lk, (unsigned long long)old, (unsigned long long)nw);/' "$REF/Raytracer/CudaComon.cuh" > "$TMP/CudaComon.cuh"
grep -q 'unsigned long long\*)lk' "$TMP/CudaComon.cuh"
NV=(nvcc -x cu -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -w -Xcompiler -fPIC -ccbin g++ -D__int32=int -D__int64=long -DNDEBUG
    -DGLM_FORCE_UNRESTRICTED_GENTYPE -I"$TMP" -I"$REF/3rdParty/glm-0.9.9.0/glm")
for src in DeviceBuffer.cpp Mesh.cpp Scene.cpp SceneTree.cpp SceneHash.cpp SceneProgressive.cpp Camera.cpp Util.cpp; do
  "${NV[@]}" -c "$TMP/$src" -o "$TMP/${src%.*}.o" &
done
"${NV[@]}" -c "$HERE/ref_headless_rt.cpp" -o "$TMP/ref_headless_rt.o" &
"${NV[@]}" -c "$HERE/ref_gpu_driver.cpp" -o "$TMP/ref_gpu_driver.o" &
"${NV[@]}" --use_fast_math -c "$TMP/BuildTree.cu" -o "$TMP/BuildTree.o" &
"${NV[@]}" --use_fast_math -c "$TMP/RTClear.cu" -o "$TMP/RTClear.o" &
wait
HOSTOBJS=("$TMP/DeviceBuffer.o" "$TMP/Mesh.o" "$TMP/Scene.o" "$TMP/SceneTree.o" "$TMP/SceneHash.o" "$TMP/SceneProgressive.o" "$TMP/Camera.o" "$TMP/Util.o"
          "$TMP/ref_headless_rt.o" "$TMP/ref_gpu_driver.o")
# (1) reference host over this repo's kernels: the bm* symbols resolve to libbeam_b200.so
nvcc -shared -ccbin g++ -o "$OUT/libbeam_refhost_b200.so" "${HOSTOBJS[@]}" -L"$(dirname "$LIBB200")" -lbeam_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../raytracercuda_b200'
# (2) reference host over the reference's own kernels
nvcc -shared -ccbin g++ -o "$OUT/libbeam_refcuda.so" "${HOSTOBJS[@]}" "$TMP/BuildTree.o" "$TMP/RTClear.o"
# (3) the reference host compiled with TREE_TYPE HASH / PROGRESSIVE (Types.h:13 patched in a scratch copy only): SceneHash.cpp and
#     SceneProgressive.cpp then bind bmResetSpace / bmInsertMeshInSpace / bmMarchSpace / bmGetCellSize / bmGetMaxHashElements and the
#     Trace2.cu size getters, all of which libbeam_b200.so exports as well.
for variant in HASH PROGRESSIVE; do
  V="$TMP/$variant"; mkdir -p "$V"
  for f in "$TMP"/*; do b="$(basename "$f")"; [ -d "$f" ] && continue; case "$b" in *.o|Types.h) continue;; esac; ln -s "$(readlink -f "$f")" "$V/$b"; done
  sed "s/^#define TREE_TYPE .*/#define TREE_TYPE          $variant/" "$REF/Raytracer/Types.h" > "$V/Types.h"
  grep -q "^#define TREE_TYPE          $variant" "$V/Types.h"
  NVV=(nvcc -x cu -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -w -Xcompiler -fPIC -ccbin g++ -D__int32=int -D__int64=long -DNDEBUG
       -DGLM_FORCE_UNRESTRICTED_GENTYPE -I"$V" -I"$REF/3rdParty/glm-0.9.9.0/glm" -include cstdio)   # SceneProgressive.cpp uses printf without <cstdio> (MSVC pulls it in)
  for src in DeviceBuffer.cpp Mesh.cpp Scene.cpp SceneTree.cpp SceneHash.cpp SceneProgressive.cpp Camera.cpp Util.cpp; do
    "${NVV[@]}" -c "$V/$src" -o "$V/${src%.*}.o" &
  done
  "${NVV[@]}" -c "$HERE/ref_headless_rt.cpp" -o "$V/ref_headless_rt.o" &
  "${NVV[@]}" -c "$HERE/ref_gpu_driver.cpp" -o "$V/ref_gpu_driver.o" &
  wait
  lower="$(echo "$variant" | tr A-Z a-z)"
  nvcc -shared -ccbin g++ -o "$OUT/libbeam_refhost_${lower}_b200.so" "$V"/*.o -L"$(dirname "$LIBB200")" -lbeam_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../raytracercuda_b200'
done
echo "built $OUT/libbeam_refhost_b200.so $OUT/libbeam_refhost_hash_b200.so $OUT/libbeam_refhost_progressive_b200.so $OUT/libbeam_refcuda.so"
