#!/bin/sh
# developer tool: build a variant of the engine library with extra -D flags for A/B timing on the GPU box
#   tools/build_variant.sh NAME [-DFOO=1 ...]   ->  citadels_self_play_b200/variants/NAME.so   (use with CTD_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../citadels_self_play_b200/csrc"
mkdir -p ../variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda \
  -Xcompiler -fPIC -shared "$@" -o ../variants/$name.so ctd_kernels.cu ctd_generic_*.cu ctd_preset_*.cu ctd_classic_*.cu
