#!/bin/bash
# Build libhudiff_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libhudiff_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden)
# HD_DEBUG_HOOKS=1: diagnostic build with the HUDIFF_DBG timing experiments and phase clocks compiled in
if [ "${HD_DEBUG_HOOKS:-0}" = "1" ]; then FLAGS+=(-DHD_DEBUG_HOOKS=1 -DHD_GEMM_CLOCKS=1 -DHD_ATTN_CLOCKS=1); fi
if [ "${HD_ATTN_CLOCKS:-0}" = "1" ]; then FLAGS+=(-DHD_ATTN_CLOCKS=1); fi
mkdir -p "$HERE/build"
pids=()
for f in gemm_tc gemm_ref attention attention_tc2 attention_tc3 attention_x3 elementwise engine; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || [ "$HERE/hd_common.cuh" -nt "$HERE/build/$f.o" ] \
     || [ "$HERE/hd_kernels.h" -nt "$HERE/build/$f.o" ] || [ "$HERE/../../include/hudiff.h" -nt "$HERE/build/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "$HERE"/build/{gemm_tc,gemm_ref,attention,attention_tc2,attention_tc3,attention_x3,elementwise,engine}.o -lcudart_static -ldl -lpthread -lrt
echo "built $OUT"
