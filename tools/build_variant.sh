#!/bin/sh
# developer tool: build a variant of the engine library with extra -D flags for A/B timing on the GPU box
#   tools/build_variant.sh NAME [-DFOO=1 ...]   ->  citadels_self_play_b200/variants/NAME.so   (use with CTD_LIB=...)
# translation units are compiled in parallel, then linked
set -e
name=$1; shift
cd "$(dirname "$0")/../citadels_self_play_b200/csrc"
mkdir -p ../variants /tmp/ctd_build_$name
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda -Xcompiler -fPIC $*"
ls ctd_kernels.cu ctd_generic_*.cu ctd_preset_*.cu ctd_classic_*.cu | xargs -P 8 -I{} sh -c "$NVCC $FLAGS -c -o /tmp/ctd_build_$name/\$(basename {} .cu).o {}"
$NVCC -shared -o ../variants/$name.so /tmp/ctd_build_$name/*.o
rm -rf /tmp/ctd_build_$name
echo built ../variants/$name.so
