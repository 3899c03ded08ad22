#!/usr/bin/env bash
# scratch/build_variant.sh NAME [extra nvcc flags]  ->  scratch/libs/libbeam_NAME.so  (select with BEAM_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch/libs
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -ccbin g++ -shared "$@" \
  -o scratch/libs/libbeam_$name.so raytracercuda_b200/csrc/{scan,sort,keygen,hierarchy,trace,bricks,rays,tree,host}.cu raytracercuda_b200/csrc/objload.cpp
