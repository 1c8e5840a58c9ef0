#!/bin/bash
# Development: build mlic_b200/libmlic_b200_<tag>.so with extra -D switches on ONE source (the other objects are reused).
#   tools/build_variant.sh <tag> <source.cu> [-DNAME=VALUE ...]     then     MLIC_LIB=mlic_b200/libmlic_b200_<tag>.so python ...
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../mlic_b200/csrc"
obj=/tmp/mlic_variant_${tag}.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 "$@" -c "$src" -o "$obj"
others=$(ls *.o | grep -v "^${src%.*}.o$")
nvcc -shared -o ../libmlic_b200_${tag}.so $obj $others -lcudart
echo ../libmlic_b200_${tag}.so
