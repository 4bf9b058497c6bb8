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

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot for two IEEE round-to-nearest operations).  The conv
// epilogues are bound by the number of instructions their few warps can issue, and half of those instructions were the
// 13 FMAs + 5 multiplies / adds of every GELU: evaluating two elements per instruction gives bit-identical results
// (same operations, same rounding, same order per element) for ~2/3 of the issue slots.
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2_pack(float a, float b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ F2 f2_pack_bits(unsigned a, unsigned b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ F2 f2_dup(float a) { return f2_pack(a, a); }
__device__ __forceinline__ void f2_unpack(F2 p, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// gelu_erf() of two values: the operation sequence of erf_branchfree() / gelu_erf() above, element for element.
__device__ __forceinline__ void gelu_erf2(float x0, float x1, float &o0, float &o1) {
  const F2 x = f2_pack(x0, x1);
  const F2 a = f2_mul(x, f2_dup(0.70710678118654752440f));
  float a0, a1;
  f2_unpack(a, a0, a1);
  const float t0 = fabsf(a0), t1 = fabsf(a1);
  const F2 t = f2_pack(t0, t1), nt = f2_pack(-t0, -t1), s = f2_mul(a, a);
  F2 r = f2_fma(f2_dup(-1.72853470e-5f), t, f2_dup(3.83197126e-4f));
  const F2 u = f2_fma(f2_dup(-3.88396438e-3f), t, f2_dup(2.42546219e-2f));
  r = f2_fma(r, s, u);
  r = f2_fma(r, t, f2_dup(-1.06777877e-1f));
  r = f2_fma(r, t, f2_dup(-6.34846687e-1f));
  r = f2_fma(r, t, f2_dup(-1.28717512e-1f));
  r = f2_fma(r, t, nt);
  float r0, r1;
  f2_unpack(r, r0, r1);
  const float big0 = copysignf(1.0f - expf(r0), a0), big1 = copysignf(1.0f - expf(r1), a1);
  F2 q = f2_fma(f2_dup(-5.96761703e-4f), s, f2_dup(4.99119423e-3f));
  q = f2_fma(q, s, f2_dup(-2.67681349e-2f));
  q = f2_fma(q, s, f2_dup(1.12819925e-1f));
  q = f2_fma(q, s, f2_dup(-3.76125336e-1f));
  q = f2_fma(q, s, f2_dup(1.28379166e-1f));
  float s0, s1;
  f2_unpack(f2_fma(q, a, a), s0, s1);
  const F2 e = f2_pack(t0 > 0.927734375f ? big0 : s0, t1 > 0.927734375f ? big1 : s1);
  f2_unpack(f2_mul(f2_mul(f2_dup(0.5f), x), f2_add(f2_dup(1.0f), e)), o0, o1);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
