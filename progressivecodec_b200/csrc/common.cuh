// Shared helpers for the pcodec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/pcodec_b200.h"

extern std::atomic<int64_t> g_pcodec_launches;

// Flight recorder: the last launches of this process (entry point, stream), so that a device fault that surfaces at a
// later synchronisation point can still be attributed (pcodec_recent_launches; printed by L.check on CUDA errors).
void pcodec_note_launch(const char *what);
#define PCODEC_COUNT_LAUNCH()                                        \
  do {                                                               \
    g_pcodec_launches.fetch_add(1, std::memory_order_relaxed);       \
    pcodec_note_launch(__func__);                                    \
  } while (0)

// Debug knob (PCODEC_SYNC_LAUNCHES=1 or pcodec_set_sync_launches(1)): synchronise the device after every launch and
// report WHICH entry point's kernel faulted (stderr + status), instead of a sticky error surfacing somewhere else later.
extern std::atomic<int> g_pcodec_sync_launches;
cudaError_t pcodec_sync_after_launch(const char *what);

// After a kernel launch: translate launch errors into the C-ABI convention (-cudaError_t).
#define PCODEC_RETURN_LAUNCH()                                                                           \
  do {                                                                                                   \
    PCODEC_COUNT_LAUNCH();                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ == cudaSuccess && g_pcodec_sync_launches.load(std::memory_order_relaxed)) e__ = pcodec_sync_after_launch(__func__); \
    return e__ == cudaSuccess ? PCODEC_OK : -(int)e__;                                                   \
  } while (0)

// Same, for entry points that counted their (several) launches themselves.
#define PCODEC_RETURN_STATUS()                                                                           \
  do {                                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ == cudaSuccess && g_pcodec_sync_launches.load(std::memory_order_relaxed)) e__ = pcodec_sync_after_launch(__func__); \
    return e__ == cudaSuccess ? PCODEC_OK : -(int)e__;                                                   \
  } while (0)

// One-time per-DEVICE opt-in (cudaFuncSetAttribute applies to the current device only): returns true when `mask` did not
// yet have this device's bit, i.e. the caller must (re)apply the attribute; call pcodec_device_mark() once it succeeded.
static inline bool pcodec_device_needs(std::atomic<uint64_t> &mask, uint64_t *bit_out) {
  int dev = 0;
  cudaGetDevice(&dev);
  *bit_out = 1ull << (dev & 63);
  return !(mask.load(std::memory_order_acquire) & *bit_out);
}

#define PCODEC_CHECK_CUDA(expr)                      \
  do {                                               \
    cudaError_t e__ = (expr);                        \
    if (e__ != cudaSuccess) return -(int)e__;        \
  } while (0)

// Experiment knobs (tiling, accumulator counts, pipeline variants) change rounding or are outright wrong-result timing
// modes: a stray environment variable on only the encoder or only the decoder host would desynchronise the entropy
// decoder silently.  They exist only in PCODEC_EXPERIMENTS builds (tools/exp_tc16.sh); release builds ignore them.
#include <stdlib.h>
static inline const char *pcodec_knob(const char *name) {
#ifdef PCODEC_EXPERIMENTS
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Branch-free fp32 erf (both polynomial branches evaluated, one selected; max error 0.97 ulp against double erf,
// checked exhaustively on a 1/97 sample of all floats).  CUDA's erff() branches on |x|, and in the conv epilogues
// neighbouring lanes sit on both sides of the threshold: the divergent regions serialised the four independent
// GELUs of a float4 (~275 clk each, measured with the clock64 timeline).  Coefficients: N. Juffa's minimax fits.
__device__ __forceinline__ float erf_branchfree(float a) {
  const float t = fabsf(a), s = a * a;
  float r = fmaf(-1.72853470e-5f, t, 3.83197126e-4f);
  const float u = fmaf(-3.88396438e-3f, t, 2.42546219e-2f);
  r = fmaf(r, s, u);
  r = fmaf(r, t, -1.06777877e-1f);
  r = fmaf(r, t, -6.34846687e-1f);
  r = fmaf(r, t, -1.28717512e-1f);
  r = fmaf(r, t, -t);
  const float big = copysignf(1.0f - expf(r), a);
  float q = -5.96761703e-4f;
  q = fmaf(q, s, 4.99119423e-3f);
  q = fmaf(q, s, -2.67681349e-2f);
  q = fmaf(q, s, 1.12819925e-1f);
  q = fmaf(q, s, -3.76125336e-1f);
  q = fmaf(q, s, 1.28379166e-1f);
  const float small = fmaf(q, a, a);
  return t > 0.927734375f ? big : small;
}

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default (exact erf form): 0.5 x (1 + erf(x / sqrt(2)))
  return 0.5f * x * (1.0f + erf_branchfree(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
