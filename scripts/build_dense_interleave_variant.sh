#!/bin/bash
# A/B build of libhelix_b200.so with the other value of HXD_INTERLEAVE (csrc/k_dense.cu) -> helix-db_b200/_variants/
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
HERE="$ROOT/helix-db_b200"
VAL="${1:-1}"
mkdir -p "$HERE/_variants"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function \
  --fmad=false -DHXD_INTERLEAVE="$VAL" -c "$HERE/csrc/k_dense.cu" -o "$HERE/_variants/k_dense_il$VAL.o"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/_variants/libhelix_b200_il$VAL.so" \
  "$HERE/_obj/hx_api.o" "$HERE/_obj/k_build.o" "$HERE/_obj/hx_shard.o" "$HERE/_variants/k_dense_il$VAL.o" -lcudart -ldl
echo "built $HERE/_variants/libhelix_b200_il$VAL.so"
