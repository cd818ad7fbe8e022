#!/bin/bash
# Build a tuning variant of the library: tools/build_variant.sh NAME "-DFOO=1 ..."  ->  build/variants/libapples_b200_NAME.so
# Run a bench against it with APPLES_B200_LIB=build/variants/libapples_b200_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
flags="$*"
src=apples_b200/csrc
out=build/variants; tmp=build/variants/obj_$name
mkdir -p $tmp
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $flags"
for f in distance select placement; do $NV -fmad=false -c $src/$f.cu -o $tmp/$f.o & done
for f in pack api dense_tc; do $NV -c $src/$f.cu -o $tmp/$f.o & done
$NV -c $src/hostio.cpp -o $tmp/hostio.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libapples_b200_$name.so $tmp/*.o
rm -rf $tmp
echo built $out/libapples_b200_$name.so
