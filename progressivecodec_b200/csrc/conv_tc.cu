// tcgen05 / TMA implicit-GEMM path (3xTF32 split accumulation) — see DESIGN.md.  Placeholder until the kernel
// lands: reports "unsupported" so pcodec_conv_taps(impl=0) routes everything to the fp32 SIMT kernel.
#include "common.cuh"

bool pcodec_conv_taps_tc_supported(const pcodec_conv_desc *) { return false; }
int pcodec_conv_taps_tc(const pcodec_conv_desc *, void *) { return PCODEC_ERR_UNSUPPORTED; }
