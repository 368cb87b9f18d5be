#!/bin/bash
# Builds the product shared library in-tree (it travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")/bwa_mem_quickassist_b200"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function,-Wno-unknown-pragmas"
mkdir -p build
for f in ksw_generic ksw_warp ksw_fast ksw_pair ksw_bin ksw_devpack ksw_global ksw_gfast ksw_align ksw_runtime; do
  if [ csrc/$f.cu -nt build/$f.o ] || [ -n "$(find csrc include ../include -newer build/$f.o 2>/dev/null | head -1)" ] || [ ! -f build/$f.o ]; then
    $NVCC $FLAGS -c csrc/$f.cu -o build/$f.o
  fi
done
if [ ! -f build/ksw_pack.o ] || [ -n "$(find csrc ../include -newer build/ksw_pack.o | head -1)" ]; then
  g++ -O3 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include -c csrc/ksw_pack.cpp -o build/ksw_pack.o
fi
if [ ! -f build/ksw_queue.o ] || [ -n "$(find csrc ../include -newer build/ksw_queue.o | head -1)" ]; then
  g++ -O2 -std=c++17 -fPIC -Wall -c csrc/ksw_queue.cpp -o build/ksw_queue.o
fi
if [ ! -f build/bwamem_ext.o ] || [ -n "$(find csrc ../include -newer build/bwamem_ext.o | head -1)" ]; then
  gcc -O2 -std=gnu99 -fPIC -Wall -c csrc/bwamem_ext.c -o build/bwamem_ext.o
fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libksw_b200.so build/ksw_generic.o build/ksw_warp.o build/ksw_fast.o build/ksw_pair.o build/ksw_bin.o build/ksw_devpack.o build/ksw_global.o build/ksw_gfast.o build/ksw_align.o build/ksw_runtime.o build/ksw_pack.o build/ksw_queue.o build/bwamem_ext.o -Xlinker -Bsymbolic-functions -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/libksw_b200.so"
