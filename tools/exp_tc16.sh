#!/bin/bash
# Experiments on the fp16 conv kernel with the PCODEC_EXPERIMENTS build (wrong results by design in the skip modes).
#   tools/exp_tc16.sh decomp | trace
cd "$(dirname "$0")/.."
cp progressivecodec_b200/libpcodec_b200.so /tmp/lib_release.so
touch progressivecodec_b200/csrc/common.cuh
PCODEC_EXPERIMENTS=1 bash progressivecodec_b200/csrc/build.sh > /tmp/build_exp.log 2>&1 || { tail /tmp/build_exp.log; echo build failed; exit 1; }
if [ "${1:-decomp}" = "trace" ]; then
  python tools/trace_tc16.py
else
  for dbg in 0 1 2 3 4 7; do
    echo "== PCODEC_TC16_DEBUG=$dbg (1 = no A loads, 2 = no B loads, 4 = hi*hi MMAs only)"
    PCODEC_TC16_DEBUG=$dbg python tools/acc_conv.py --impls 3 --batch 32 2>&1 | grep -E "512->|224->|192-> 192 @256|128->  64"
  done
fi
cp /tmp/lib_release.so progressivecodec_b200/libpcodec_b200.so
touch progressivecodec_b200/csrc/common.cuh
