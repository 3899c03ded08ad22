#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference (d1kkop/RaytracerCuda, "Beam") in its own serial
# host-emulation mode into oracle/_ref/ so the CPU restatement (oracle/oracle.cpp) can be pinned against it.
#
#  * Sources are compiled where they lie under $REF (default /root/reference); nothing is copied into the repo.
#    The single change the CPU mode needs, `#define CUDA 1 -> 0` (Raytracer/Types.h:8), is applied to a scratch
#    copy of that one header in a mktemp directory that also holds symlinks to the other sources (quote-includes
#    resolve next to the including file, so the patched header has to sit beside them). The scratch dir is removed.
#  * Only outputs land in oracle/_ref/ (git-ignored, not gpurun-ignored): libbeam_ref.so, libbeam_ref_hash.so.
#  * Flags follow SURVEY.md section 8(c): g++ -std=c++17 -O2 -DNDEBUG -ffp-contract=off, MSVC integer names mapped.
#  * The reference's own build system (VS2017 .vcxproj) is not used.
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/Raytracer" ]; then echo "build_ref.sh: $REF/Raytracer not present; keeping prebuilt oracle/_ref if any"; exit 0; fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
for f in "$REF"/Raytracer/*; do
  b="$(basename "$f")"
  [ "$b" = "Types.h" ] && continue
  ln -s "$f" "$TMP/$b"
done
sed 's/^#define CUDA 1/#define CUDA 0/' "$REF/Raytracer/Types.h" > "$TMP/Types.h"
grep -q '^#define CUDA 0' "$TMP/Types.h"
CXX="${ORACLE_CXX:-g++}"   # not $CXX: this image exports a CXX with a static libstdc++, which breaks inside python
FLAGS=(-std=c++17 -O2 -DNDEBUG -fPIC -ffp-contract=off -w -D__int32=int -D__int64=long -DGLM_FORCE_UNRESTRICTED_GENTYPE
       -I"$TMP" -I"$REF/3rdParty/glm-0.9.9.0/glm" -I/usr/local/cuda/include)
OBJS=()
for src in BuildTree.cu RTClear.cu DeviceBuffer.cpp Mesh.cpp Scene.cpp SceneTree.cpp SceneHash.cpp SceneProgressive.cpp Camera.cpp Util.cpp; do
  o="$TMP/${src%.*}.o"
  "$CXX" "${FLAGS[@]}" -x c++ -c "$TMP/$src" -o "$o" &
  OBJS+=("$o")
done
"$CXX" "${FLAGS[@]}" -c "$HERE/ref_headless_rt.cpp" -o "$TMP/ref_headless_rt.o" &
"$CXX" "${FLAGS[@]}" -c "$HERE/ref_driver.cpp" -o "$TMP/ref_driver.o" &
"$CXX" "${FLAGS[@]}" -x c++ -c "$TMP/Hash.cu" -o "$TMP/Hash.o" &
"$CXX" "${FLAGS[@]}" -c "$HERE/ref_hash_driver.cpp" -o "$TMP/ref_hash_driver.o" &
wait
"$CXX" -shared -o "$OUT/libbeam_ref.so" "${OBJS[@]}" "$TMP/ref_headless_rt.o" "$TMP/ref_driver.o" -Wl,--no-as-needed -lstdc++ -lm
"$CXX" -shared -o "$OUT/libbeam_ref_hash.so" "$TMP/Hash.o" "$TMP/ref_hash_driver.o" -Wl,--no-as-needed -lstdc++ -lm
echo "built $OUT/libbeam_ref.so $OUT/libbeam_ref_hash.so"
