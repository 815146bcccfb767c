#!/bin/bash
# A/B aid: build a variant of the library with extra -D flags for ens_tc.cu into tools/lib_<tag>.so
# usage: bash tools/build_variant.sh <tag> -DFOO [-DBAR ...]     (run build.sh first: the other objects are reused)
set -e
cd "$(dirname "$0")/.."
TAG=$1; shift
PKG="constrained-model-based-policy-optimization_b200"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
$NVCC $FLAGS "$@" -c $PKG/csrc/ens_tc.cu -o build/ens_tc_$TAG.o
objs=""
for f in capi gae ens_f32 rollout policy_pack archive train; do objs="$objs build/$f.o"; done
$NVCC -shared -o tools/lib_$TAG.so $objs build/ens_tc_$TAG.o -cudart static -lcublas -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo "built tools/lib_$TAG.so"
