#!/bin/bash
# usage: scripts/build_variant.sh <tag> <file.cu> <nvcc -D flags...>  -- libchefsi_b200_<tag>.so with ONE translation unit rebuilt
# with extra flags (A/B experiments; select with CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_<tag>.so)
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../sparc_b200/csrc"
make -s -j8
mkdir -p build_$tag
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c $src -o build_$tag/${src%.cu}.o 2> build_$tag/${src%.cu}.ptxas.log
objs=$(ls build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libchefsi_b200_$tag.so $objs build_$tag/${src%.cu}.o -cudart static -ldl -lpthread
grep -E "spill|Used" build_$tag/${src%.cu}.ptxas.log | paste - - | sort | uniq -c | head -8
