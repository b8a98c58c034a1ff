#!/bin/bash
# usage: tools/build_variant.sh NAME file.cu -DFLAG=... ; builds ihmr_b200/_lib/variants/libihmr_NAME.so
# (the named source recompiled with the extra flags, linked with the objects of the regular build)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p ihmr_b200/_lib/variants
obj=ihmr_b200/_lib/variants/${name}_${src%.cu}.o
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -O2 \
  --expt-relaxed-constexpr "$@" -c ihmr_b200/csrc/$src -o $obj
others=$(ls ihmr_b200/_lib/*.o | grep -v "/${src%.cu}.o" | grep -v "/[a-z0-9]*_${src%.cu}.o")
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a $obj $others -o ihmr_b200/_lib/variants/libihmr_${name}.so
echo ihmr_b200/_lib/variants/libihmr_${name}.so
