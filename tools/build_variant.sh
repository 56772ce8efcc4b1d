#!/bin/bash
# usage: tools/build_variant.sh NAME OBJECT_STEM PART "extra nvcc flags"   one object rebuilt with extra flags, linked with the
# other current objects into tools/_probe/var/lib_NAME.so (for A/B runs with tools/_probe/pack_probe)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_probe/var
C=vc2_reference_b200/csrc
src=$C/$2.cu; [ "$2" = dwt_tile_inv ] && src=$C/dwt_tile.cu; [ "$2" = dwt_tile_fwd ] && src=$C/dwt_tile.cu
nvcc -O3 -std=c++17 -lineinfo -diag-suppress 177 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC ${3:+-DVC2_DWT_PART=$3} $4 -c $src -o tools/_probe/var/$1_$2.o
objs=""
for o in dwt_fwd dwt_inv dwt_tile_fwd dwt_tile_inv slices cabi; do
  if [ "$o" = "$2" ]; then objs="$objs tools/_probe/var/$1_$2.o"; else objs="$objs $C/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_probe/var/lib_$1.so $objs -lcudart
echo built tools/_probe/var/lib_$1.so
