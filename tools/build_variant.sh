#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. -DFLAG2=.." : builds spectral_b200/lib_NAME/libspectral.so with extra nvcc defines
# (A/B experiments on the GPU box: SPECTRAL_LIB_DIR=spectral_b200/lib_NAME python bench.py ...)
set -e
cd "$(dirname "$0")/../spectral_b200/csrc"
OUT=../lib_$1
mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
$NV $2 -Xptxas -v -c capi.cu -o $OUT/capi.o 2> $OUT/ptxas_capi.log
$NV --fmad=false -c corridor.cu -o $OUT/corridor.o
$NV -shared -o $OUT/libspectral.so $OUT/capi.o $OUT/corridor.o -cudart shared
grep -A3 "k_qpaILi" $OUT/ptxas_capi.log | grep -E "Used|spill" | head -8
