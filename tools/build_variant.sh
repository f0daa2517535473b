#!/bin/bash
# usage: tools/build_variant.sh <name> [-DMACRO ...]   -> variants/<name>.so (development aid for A/B runs on the GPU box)
set -e
cd "$(dirname "$0")/../srslte-emane_b200"
name=$1; shift
mkdir -p ../variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" \
  csrc/tdec_kernels.cu csrc/frontend_kernels.cu csrc/capi.cu csrc/lte_tables.cpp -o ../variants/$name.so
