// Batched, bit-exact rANS coder: one warp per independent stream.
//
// Replaces compressai.ans (reference: compress/cpp_exts/rans/rans_interface.cpp:99-275).  A stream is one
// (image, slice) symbol plane; its bytes are identical to the reference encoder's, so streams are
// interchangeable with the CPU coder in both directions.  A single stream is an inherently serial state
// chain; parallelism comes from (a) many streams per launch and (b) using the 32 lanes of the warp to take
// everything that does NOT depend on the state off the chain: coalesced symbol/index loads, table
// look-ups, escape classification (encoder) and the CDF search (decoder: 32 CDF entries are compared per
// ballot instead of the reference's linear std::find_if).
#include "common.cuh"
#include "rans_core.h"

using namespace pcodec;

namespace {

constexpr int kWarpsPerBlock = 1;  // one stream per CTA keeps few streams spread over many SMs

// ------------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
rans_encode_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int n_streams,
                   int64_t n, const int32_t *__restrict__ cdfs, int cdf_stride,
                   const int32_t *__restrict__ cdf_sizes, const int32_t *__restrict__ offsets,
                   uint32_t *__restrict__ scratch, int64_t scratch_words, int32_t *__restrict__ n_words_out,
                   int32_t *__restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= n_streams) return;
  const int32_t *sym = symbols + (int64_t)s * n;
  const int32_t *idx = indexes + (int64_t)s * n;
  uint32_t *out = scratch + (int64_t)s * scratch_words;
  int64_t wpos = scratch_words;  // next free slot is wpos-1 (stream grows downwards)
  bool overflow = false;
  uint64_t x = kRansLower;

  auto emit = [&](uint32_t word) {
    if (wpos > 0) {
      --wpos;
      if (lane == 0) out[wpos] = word;
    } else {
      overflow = true;
    }
  };

  const int64_t n_chunks = (n + 31) / 32;
  for (int64_t c = n_chunks - 1; c >= 0; --c) {
    const int64_t i = c * 32 + lane;
    const bool valid = i < n;
    uint32_t packed = 0, raw = 0;
    bool esc = false;
    if (valid) {
      const int32_t t = __ldg(idx + i);
      const int32_t sy = __ldg(sym + i);
      const int32_t maxv = __ldg(cdf_sizes + t) - 2;
      int32_t slot;
      classify(sy, __ldg(offsets + t), maxv, slot, raw, esc);
      const int32_t *row = cdfs + (int64_t)t * cdf_stride;
      const uint32_t start = (uint32_t)__ldg(row + slot);
      const uint32_t freq = ((uint32_t)__ldg(row + slot + 1) - start) & 0xFFFFu;
      packed = (start & 0xFFFFu) | (freq << 16);
    }
    const uint32_t esc_mask = __ballot_sync(0xFFFFFFFFu, esc);
    const int last = (int)min((int64_t)32, n - c * 32) - 1;
    for (int j = last; j >= 0; --j) {
      uint32_t word;
      if ((esc_mask >> j) & 1u) {  // warp-uniform branch
        const uint32_t r = __shfl_sync(0xFFFFFFFFu, raw, j);
        const int nb = nibble_count(r);
        for (int k = nb - 1; k >= 0; --k)
          if (enc_put_bits4(x, (r >> (k * kBypassBits)) & kBypassMax, word)) emit(word);
        if (enc_put_bits4(x, (uint32_t)(nb % (int)kBypassMax), word)) emit(word);
        for (int k = 0; k < nb / (int)kBypassMax; ++k)
          if (enc_put_bits4(x, kBypassMax, word)) emit(word);
      }
      const uint32_t p = __shfl_sync(0xFFFFFFFFu, packed, j);
      if (enc_put(x, p & 0xFFFFu, p >> 16, word)) emit(word);
    }
  }
  // flush: low word first in memory (rans64.h:96-103)
  emit((uint32_t)(x >> 32));
  emit((uint32_t)x);
  if (lane == 0) {
    n_words_out[s] = (int32_t)(scratch_words - wpos);
    if (overflow) atomicExch(status, PCODEC_ERR_OVERFLOW);
  }
}

// exclusive scan of per-stream byte counts -> out_offsets (single CTA; n_streams is small)
__global__ void __launch_bounds__(1024)
rans_offsets_kernel(const int32_t *__restrict__ n_words, int n_streams, int64_t *__restrict__ out_offsets,
                    int64_t out_cap, int32_t *__restrict__ status) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_streams; base += 1024) {
    const int i = base + tid;
    int64_t v = i < n_streams ? 4ll * n_words[i] : 0;
    int64_t incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      int64_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_sums[lane];
      int64_t wi = w;
      for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xFFFFFFFFu, wi, d);
        if (lane >= d) wi += o;
      }
      warp_sums[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t excl = carry + warp_sums[warp] + incl - v;
    if (i < n_streams) out_offsets[i] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) {
    out_offsets[n_streams] = carry_s;
    if (carry_s > out_cap) atomicExch(status, PCODEC_ERR_OVERFLOW);
  }
}

__global__ void __launch_bounds__(256)
rans_compact_kernel(const uint32_t *__restrict__ scratch, int64_t scratch_words, const int32_t *__restrict__ n_words,
                    const int64_t *__restrict__ out_offsets, int n_streams, uint8_t *__restrict__ out_bytes,
                    int64_t out_cap) {
  const int s = blockIdx.x;
  if (out_offsets[n_streams] > out_cap) return;  // flagged by the scan kernel
  const int32_t nw = n_words[s];
  const uint32_t *src = scratch + (int64_t)s * scratch_words + (scratch_words - nw);
  uint32_t *dst = reinterpret_cast<uint32_t *>(out_bytes + out_offsets[s]);
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < nw; i += blockDim.x * gridDim.y) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------------
struct WordReader {
  const uint32_t *base;
  int64_t n_words;
  int64_t rp;          // next word to read
  int64_t cache_base;  // lane l caches word[cache_base + l]
  uint32_t cache;
  int lane;
  __device__ __forceinline__ void fill(int64_t from) {
    cache_base = from;
    const int64_t i = from + lane;
    cache = i < n_words ? __ldg(base + i) : 0u;
  }
  __device__ __forceinline__ uint32_t next() {
    if (rp - cache_base >= 32) fill(rp);
    const uint32_t w = __shfl_sync(0xFFFFFFFFu, cache, (int)(rp - cache_base));
    ++rp;
    return w;
  }
};

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
rans_decode_kernel(const uint8_t *__restrict__ in_bytes, const int64_t *__restrict__ in_offsets, int n_streams,
                   int64_t n, const int32_t *__restrict__ indexes, const int32_t *__restrict__ cdfs, int cdf_stride,
                   const int32_t *__restrict__ cdf_sizes, const int32_t *__restrict__ offsets,
                   int32_t *__restrict__ out_symbols) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= n_streams) return;
  const int32_t *idx = indexes + (int64_t)s * n;
  int32_t *out = out_symbols + (int64_t)s * n;
  WordReader rd;
  rd.base = reinterpret_cast<const uint32_t *>(in_bytes + in_offsets[s]);
  rd.n_words = (in_offsets[s + 1] - in_offsets[s]) / 4;
  rd.lane = lane;
  rd.rp = 0;
  rd.fill(0);
  uint64_t x = (uint64_t)rd.next();
  x |= (uint64_t)rd.next() << 32;

  const int64_t n_chunks = (n + 31) / 32;
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t i = c * 32 + lane;
    int32_t t_l = 0, size_l = 2, off_l = 0;
    if (i < n) {
      t_l = __ldg(idx + i);
      size_l = __ldg(cdf_sizes + t_l);
      off_l = __ldg(offsets + t_l);
    }
    int32_t my_value = 0;
    const int count = (int)min((int64_t)32, n - c * 32);
    for (int j = 0; j < count; ++j) {
      const int32_t t = __shfl_sync(0xFFFFFFFFu, t_l, j);
      const int32_t size = __shfl_sync(0xFFFFFFFFu, size_l, j);
      const int32_t off = __shfl_sync(0xFFFFFFFFu, off_l, j);
      const int32_t *row = cdfs + (int64_t)t * cdf_stride;
      const uint32_t cf = (uint32_t)(x & 0xFFFFu);
      // 32-wide probe window centred on the slot of symbol value 0 (slot = -offset)
      int32_t ws = -off - 15;
      ws = max(0, min(ws, size - 32));
      const int32_t pos = ws + lane;
      const uint32_t val = pos < size ? (uint32_t)__ldg(row + pos) : 0xFFFFFFFFu;
      const uint32_t gt = __ballot_sync(0xFFFFFFFFu, val > cf);
      int32_t slot;
      uint32_t start, next;
      const int first = __ffs(gt) - 1;  // -1 when no probe entry exceeds cf
      if (first > 0 || (first == 0 && ws == 0)) {
        // first == 0 with ws == 0 cannot happen for a valid CDF (cdf[0] = 0 <= cf); clamped for safety
        slot = max(ws + first - 1, 0);
        start = __shfl_sync(0xFFFFFFFFu, val, max(first - 1, 0));
        next = __shfl_sync(0xFFFFFFFFu, val, first);
      } else {
        // outside the window: uniform binary search for the last entry <= cf
        int32_t lo = (first == 0) ? 0 : ws + 31;       // cdf[lo] <= cf
        int32_t hi = (first == 0) ? ws : size - 1;     // cdf[hi] > cf
        while (hi - lo > 1) {
          const int32_t mid = (lo + hi) >> 1;
          if ((uint32_t)__ldg(row + mid) <= cf) lo = mid; else hi = mid;
        }
        slot = lo;
        start = (uint32_t)__ldg(row + lo);
        next = (uint32_t)__ldg(row + lo + 1);
      }
      if (dec_advance(x, start, next - start)) x = (x << 32) | rd.next();
      int32_t value = slot;
      const int32_t maxv = size - 2;
      if (slot == maxv) {  // escape (warp-uniform)
        uint32_t v;
        if (dec_get_bits4(x, v)) x = (x << 32) | rd.next();
        int32_t nb = (int32_t)v;
        while (v == kBypassMax) {
          if (dec_get_bits4(x, v)) x = (x << 32) | rd.next();
          nb += (int32_t)v;
        }
        uint32_t raw = 0;
        for (int k = 0; k < nb; ++k) {
          if (dec_get_bits4(x, v)) x = (x << 32) | rd.next();
          if (k < 8) raw |= v << (k * kBypassBits);
        }
        value = (int32_t)(raw >> 1);
        if (raw & 1u) value = -value - 1; else value += maxv;
      }
      if (lane == j) my_value = value + off;
    }
    if (i < n) out[i] = my_value;
  }
}

}  // namespace

extern "C" int pcodec_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int n_streams,
                                        int64_t n_per_stream, const int32_t *cdfs, int cdf_stride,
                                        const int32_t *cdf_sizes, const int32_t *offsets, int n_tables,
                                        uint32_t *scratch, int64_t scratch_words, int32_t *n_words,
                                        uint8_t *out_bytes, int64_t out_cap, int64_t *out_offsets, int32_t *status,
                                        void *stream) {
  (void)n_tables;
  if (n_streams <= 0 || n_per_stream < 0 || scratch_words < 2 || !scratch || !n_words || !out_bytes || !out_offsets ||
      !status || (n_per_stream > 0 && (!symbols || !indexes)))
    return PCODEC_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  PCODEC_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  const int blocks = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
  rans_encode_kernel<<<blocks, 32 * kWarpsPerBlock, 0, st>>>(symbols, indexes, n_streams, n_per_stream, cdfs,
                                                              cdf_stride, cdf_sizes, offsets, scratch, scratch_words,
                                                              n_words, status);
  PCODEC_COUNT_LAUNCH();
  rans_offsets_kernel<<<1, 1024, 0, st>>>(n_words, n_streams, out_offsets, out_cap, status);
  PCODEC_COUNT_LAUNCH();
  rans_compact_kernel<<<dim3(n_streams, 4), 256, 0, st>>>(scratch, scratch_words, n_words, out_offsets, n_streams,
                                                          out_bytes, out_cap);
  PCODEC_COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PCODEC_OK : -(int)e;
}

extern "C" int pcodec_rans_decode_batch(const uint8_t *in_bytes, const int64_t *in_offsets, int n_streams,
                                        int64_t n_per_stream, const int32_t *indexes, const int32_t *cdfs,
                                        int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int n_tables,
                                        int32_t *out_symbols, void *stream) {
  (void)n_tables;
  if (n_streams <= 0 || n_per_stream < 0 || !in_bytes || !in_offsets) return PCODEC_ERR_BAD_ARG;
  if (n_per_stream == 0) return PCODEC_OK;
  if (!indexes || !out_symbols) return PCODEC_ERR_BAD_ARG;
  const int blocks = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
  rans_decode_kernel<<<blocks, 32 * kWarpsPerBlock, 0, as_stream(stream)>>>(
      in_bytes, in_offsets, n_streams, n_per_stream, indexes, cdfs, cdf_stride, cdf_sizes, offsets, out_symbols);
  PCODEC_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// HOST self-test of rans_core.h (same inline arithmetic as the kernels, scalar walk).  Used only by the
// CPU unit tests to pin divmod_u63_u16 / enc_put / dec_advance against the oracle without a GPU; the
// product path never calls it.
// ------------------------------------------------------------------------------------------------
extern "C" int64_t pcodec_selftest_rans_core_encode(const int32_t *symbols, const int32_t *indexes, int64_t n,
                                                    const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                                    const int32_t *offsets, uint32_t *words, int64_t cap_words) {
  int64_t wpos = cap_words;
  uint64_t x = kRansLower;
  uint32_t word;
  for (int64_t i = n - 1; i >= 0; --i) {
    const int32_t t = indexes[i];
    const int32_t maxv = cdf_sizes[t] - 2;
    int32_t slot;
    uint32_t raw;
    bool esc;
    classify(symbols[i], offsets[t], maxv, slot, raw, esc);
    const int32_t *row = cdfs + (int64_t)t * cdf_stride;
    if (esc) {
      const int nb = nibble_count(raw);
      for (int k = nb - 1; k >= 0; --k)
        if (enc_put_bits4(x, (raw >> (k * kBypassBits)) & kBypassMax, word)) { if (wpos <= 0) return -1; words[--wpos] = word; }
      if (enc_put_bits4(x, (uint32_t)(nb % (int)kBypassMax), word)) { if (wpos <= 0) return -1; words[--wpos] = word; }
      for (int k = 0; k < nb / (int)kBypassMax; ++k)
        if (enc_put_bits4(x, kBypassMax, word)) { if (wpos <= 0) return -1; words[--wpos] = word; }
    }
    const uint32_t start = (uint32_t)row[slot];
    const uint32_t freq = ((uint32_t)row[slot + 1] - start) & 0xFFFFu;
    if (enc_put(x, start & 0xFFFFu, freq, word)) { if (wpos <= 0) return -1; words[--wpos] = word; }
  }
  if (wpos < 2) return -1;
  words[--wpos] = (uint32_t)(x >> 32);
  words[--wpos] = (uint32_t)x;
  return cap_words - wpos;  // words used, stored at the END of `words`
}
