#!/bin/bash
# Builds measurement variants of libartes_gpu.so (one optimisation of engine2.cuh switched off each) into build/variants/.
# bench.py picks a variant through ARTES_GPU_LIB.  usage: tools/build_variants.sh name:"-DFLAG=0 ..." ...
set -e
cd "$(dirname "$0")/../artes_b200/csrc"
mkdir -p ../../build/variants
make -s transport_faithful.o fma_peak.o ingest.o artes_gpu.o
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -fmad=true $flags \
      -Xptxas -v -c transport_fast.cu -o ../../build/variants/fast_$name.o 2> ../../build/variants/fast_$name.ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../build/variants/libartes_gpu_$name.so \
      transport_faithful.o ../../build/variants/fast_$name.o fma_peak.o ingest.o artes_gpu.o -ldl
  echo "built $name ($flags): $(grep -A2 'transport3_kernelILi256ELi512ELi2ELb0ELb0ELb0ELb0' ../../build/variants/fast_$name.ptxas.log | grep spill)"
done
