#!/bin/bash
# dev: build a variant of the library with extra -D flags for ONE source file
#   tools/build_variant.sh NAME file.cu "-DFLAG1 -DFLAG2"   ->  gpurun_variants/libdgnn_NAME.so
set -e
cd "$(dirname "$0")/../dgnn_b200/csrc"
name=$1; src=$2; flags=$3
mkdir -p ../../gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $src -o /tmp/variant_$name.o
objs=""
for f in head graph layer_fwd layer_bwd layer_tc dw_tc gather_tc upd sampler graphcut; do
  if [ "$f.cu" == "$src" ]; then objs="$objs /tmp/variant_$name.o"; else objs="$objs $f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libdgnn_$name.so $objs -lcudart
echo built gpurun_variants/libdgnn_$name.so
