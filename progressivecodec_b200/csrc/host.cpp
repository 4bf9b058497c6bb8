// Host-side pieces of the C-ABI: version / device info / launch counter and pmf_to_quantized_cdf.
#include <cuda_runtime.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

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

std::atomic<int> g_pcodec_sync_launches{[] {
  const char *e = getenv("PCODEC_SYNC_LAUNCHES");
  return e && atoi(e) != 0 ? 1 : 0;
}()};

cudaError_t pcodec_sync_after_launch(const char *what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) fprintf(stderr, "[pcodec] device fault after %s: %s (%d)\n", what, cudaGetErrorString(e), (int)e);
  return e;
}

namespace {
constexpr int kRing = 64;
struct Note { const char *what; unsigned long long seq; unsigned long tid; };
Note g_ring[kRing];
std::atomic<unsigned long long> g_ring_seq{0};
}  // namespace

void pcodec_note_launch(const char *what) {
  const unsigned long long seq = g_ring_seq.fetch_add(1, std::memory_order_relaxed);
  Note &n = g_ring[seq % kRing];
  n.what = what;
  n.seq = seq;
  n.tid = (unsigned long)pthread_self();
}

extern "C" int pcodec_recent_launches(char *buf, int cap) {
  if (!buf || cap <= 0) return 0;
  int pos = 0;
  const unsigned long long end = g_ring_seq.load();
  const unsigned long long begin = end > kRing ? end - kRing : 0;
  for (unsigned long long q = begin; q < end && pos < cap - 96; ++q) {
    const Note n = g_ring[q % kRing];
    pos += snprintf(buf + pos, cap - pos, "#%llu thread %lx %s\n", n.seq, n.tid, n.what ? n.what : "?");
  }
  buf[pos < cap ? pos : cap - 1] = 0;
  return pos;
}

extern "C" void pcodec_set_sync_launches(int on) { g_pcodec_sync_launches.store(on ? 1 : 0); }

extern "C" const char *pcodec_error_string(int status) {
  switch (status) {
    case PCODEC_OK: return "ok";
    case PCODEC_ERR_BAD_ARG: return "bad argument";
    case PCODEC_ERR_UNSUPPORTED: return "unsupported configuration";
    case PCODEC_ERR_OVERFLOW: return "capacity overflow";
    default: return status < 0 ? cudaGetErrorString((cudaError_t)(-status)) : "unknown status";
  }
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
