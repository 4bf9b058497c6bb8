// Host-side pieces of the C-ABI: version / device info / launch counter and pmf_to_quantized_cdf.
#include <cuda_runtime.h>
#include <math.h>

#include <atomic>
#include <vector>

#include "../../include/pcodec_b200.h"

std::atomic<int64_t> g_pcodec_launches{0};

extern "C" int pcodec_version(void) { return 100; }

extern "C" int pcodec_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -(int)e;
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return -(int)e;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return PCODEC_OK;
}

extern "C" int64_t pcodec_launch_count(void) { return g_pcodec_launches.load(); }
extern "C" void pcodec_reset_launch_count(void) { g_pcodec_launches.store(0); }

// pmf -> quantised CDF (reference: compress/cpp_exts/ops/ops.cpp:10-67).  Setup-time (update()), host only.
// Differences from the reference: returns an error instead of asserting when nothing can be stolen, and
// keeps per-bin frequencies so a steal is an O(n) scan only for the (rare) zero-width bins.
extern "C" int pcodec_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  if (!pmf || !cdf || n <= 0 || precision < 1 || precision > 16) return PCODEC_ERR_BAD_ARG;
  const uint32_t one = 1u << precision;
  std::vector<uint32_t> f((size_t)n + 1);
  f[0] = 0;
  uint32_t total = 0;
  for (int i = 0; i < n; ++i) {
    f[i + 1] = (uint32_t)roundf(pmf[i] * (float)one);
    total += f[i + 1];
  }
  if (total == 0) return PCODEC_ERR_BAD_ARG;
  uint32_t run = 0;
  for (int i = 0; i <= n; ++i) {
    run += (uint32_t)(((uint64_t)one * f[i]) / total);
    cdf[i] = run;
  }
  cdf[n] = one;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best_freq = ~0u;
    int best = -1;
    for (int j = 0; j < n; ++j) {
      const uint32_t fr = cdf[j + 1] - cdf[j];
      if (fr > 1 && fr < best_freq) {
        best_freq = fr;
        best = j;
      }
    }
    if (best < 0) return PCODEC_ERR_BAD_ARG;
    if (best < i) {
      for (int j = best + 1; j <= i; ++j) cdf[j]--;
    } else {
      for (int j = i + 1; j <= best; ++j) cdf[j]++;
    }
  }
  return PCODEC_OK;
}
