#!/bin/bash
# Development helper: builds libksw_b200.so with extra -D flags for the extension kernels into
# bwa_mem_quickassist_b200/build/variants/libksw_b200_<name>.so (select it with KSW_B200_LIB=<path>;
# scripts/ab_kernel.py times every variant).   scripts/build_variant.sh <name> [-DFLAG ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../bwa_mem_quickassist_b200"
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3"
nvcc $FLAGS "$@" -c csrc/ksw_fast.cu -o build/variants/ksw_fast_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libksw_b200_$name.so build/ksw_generic.o build/ksw_warp.o build/variants/ksw_fast_$name.o build/ksw_pair.o build/ksw_bin.o build/ksw_devpack.o build/ksw_global.o build/ksw_gfast.o build/ksw_align.o build/ksw_runtime.o build/ksw_pack.o build/ksw_queue.o build/bwamem_ext.o -Xlinker -Bsymbolic-functions -lcudart_static -lpthread -ldl -lrt
echo build/variants/libksw_b200_$name.so
