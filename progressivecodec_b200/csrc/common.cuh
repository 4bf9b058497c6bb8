// Shared helpers for the pcodec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/pcodec_b200.h"

extern std::atomic<int64_t> g_pcodec_launches;

#define PCODEC_COUNT_LAUNCH() g_pcodec_launches.fetch_add(1, std::memory_order_relaxed)

// After a kernel launch: translate launch errors into the C-ABI convention (-cudaError_t).
#define PCODEC_RETURN_LAUNCH()                        \
  do {                                                \
    PCODEC_COUNT_LAUNCH();                            \
    cudaError_t e__ = cudaGetLastError();             \
    return e__ == cudaSuccess ? PCODEC_OK : -(int)e__; \
  } while (0)

#define PCODEC_CHECK_CUDA(expr)                      \
  do {                                               \
    cudaError_t e__ = (expr);                        \
    if (e__ != cudaSuccess) return -(int)e__;        \
  } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default (exact erf form): 0.5 x (1 + erf(x / sqrt(2)))
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
