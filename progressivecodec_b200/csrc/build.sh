#!/bin/bash
# Build libpcodec_b200.so for sm_100a (in-tree; travels to the GPU box with the repo snapshot).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libpcodec_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${PCODEC_EXPERIMENTS:+-DPCODEC_EXPERIMENTS} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
OBJS=()
for f in rans.cu entropy_ops.cu conv_simt.cu conv_tc.cu conv_tc16.cu attention.cu host.cpp; do
  o=build/${f%.*}.o
  mkdir -p build
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ rans_core.h -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ ../../include/pcodec_b200.h -nt "$o" ]; then
    echo "[nvcc] $f"
    $NVCC $FLAGS ${PCODEC_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o"
  fi
  OBJS+=("$o")
done
$NVCC -shared -o $OUT "${OBJS[@]}" -lcudart -lcuda
echo "built $OUT"
