#!/bin/bash
# tools/nvcc_alpaka.sh <src> <out> [extra flags] -- the build line for a translation unit using include/alpaka
SRC=$1; OUT=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
exec /usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -std=c++20 -O3 -lineinfo --expt-relaxed-constexpr --extended-lambda \
  -gencode arch=compute_100a,code=sm_100a -cudart shared -x cu -I"$ROOT/include" "$@" "$SRC" -o "$OUT" \
  -L"$ROOT/alpaka_b200/lib" -lalpaka_b200 -Xlinker -rpath,"$ROOT/alpaka_b200/lib" -Xlinker -rpath,/usr/local/cuda/lib64
