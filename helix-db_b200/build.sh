#!/bin/bash
# Builds helix-db_b200/libhelix_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/libhelix_b200.so"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function
       --fmad=false -Xptxas -v)
SRCS=("$HERE/csrc/hx_api.cu" "$HERE/csrc/k_build.cu" "$HERE/csrc/k_dense.cu")
mkdir -p "$HERE/_obj"
OBJS=()
for s in "${SRCS[@]}"; do
  o="$HERE/_obj/$(basename "${s%.cu}").o"
  if [[ ! -f "$o" || "$s" -nt "$o" || -n "$(find "$HERE/csrc" "$HERE/../include" -newer "$o" \( -name '*.cuh' -o -name '*.hpp' -o -name '*.h' \) -print -quit)" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" 2> "$o.log" || { cat "$o.log"; exit 1; }
  fi
  OBJS+=("$o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${OBJS[@]}" -lcudart
echo "built $OUT"
