#!/bin/bash
# Diagnostic: build libfrecsys_b200_<tag>.so with extra -D flags on ONE source file (timing experiments only).
#   tools/build_variant.sh <tag> <file.cu> -DFLAG ...
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../safer2-recommender_b200/csrc"
make -s -j8
NVCC=/usr/local/cuda/bin/nvcc
$NVCC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DFRX_WITH_NCCL "$@" -c $src -o /tmp/${src%.cu}_$tag.o
objs=""
for f in *.o; do if [ "$f" = "${src%.cu}.o" ]; then objs="$objs /tmp/${src%.cu}_$tag.o"; else objs="$objs $f"; fi; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libfrecsys_b200_$tag.so $objs -lnccl
echo built ../libfrecsys_b200_$tag.so
