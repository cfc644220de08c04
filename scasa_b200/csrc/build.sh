#!/bin/sh
# Builds libscasa_cuda.so in-tree for sm_100a (CUDA 12.9).  Static cudart so the library can be
# loaded next to torch's own runtime (or by R) without version coupling.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC,-O2,-Wall -Xptxas -v --shared -cudart static \
      -I../../include -o ../libscasa_cuda.so scasa_cuda.cu scasa_host.cpp scasa_bfh.cpp scasa_bus.cpp scasa_write.cpp 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
