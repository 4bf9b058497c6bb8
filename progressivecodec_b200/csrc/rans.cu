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
// Division-free state update (Alverson reciprocals, as rans64.h's Rans64EncSymbol does on the CPU): for
// 2 <= freq < 2^16, shift = ceil(log2 freq), rcp = ceil(2^(shift+63) / freq); then for every x < 2^63
//   x / freq == mulhi64(x, rcp) >> (shift - 1)           (exact)
// and C(s,x) = ((x/freq) << 16) + x % freq + start = x + start + (x/freq) * (65536 - freq).
// The 64K-entry reciprocal table lives in global memory (512 KB, L2 resident) and is filled once per process.
__device__ uint64_t g_rcp_table[65536];

__global__ void rans_build_rcp_kernel() {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= 65536) return;
  g_rcp_table[f] = rcp_for(f);
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
rans_encode_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int n_streams,
                   int64_t n, const int32_t *__restrict__ cdfs, int cdf_stride,
                   const int32_t *__restrict__ cdf_sizes, const int32_t *__restrict__ offsets,
                   uint32_t *__restrict__ scratch, int64_t scratch_words, int32_t *__restrict__ n_words_out,
                   int32_t *__restrict__ status, const int64_t *__restrict__ seg_start,
                   const int32_t *__restrict__ seg_count) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (s >= n_streams) return;
  // segment mode: stream s = elements [seg_start[s], seg_start[s] + seg_count[s]) of the flat symbol / index arrays
  const int64_t first = seg_start ? seg_start[s] : (int64_t)s * n;
  if (seg_count) n = seg_count[s];
  const int32_t *sym = symbols + first;
  const int32_t *idx = indexes + first;
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
  // software pipeline: the symbol/index/table loads of chunk c-1 are issued before the serial walk of chunk c
  uint32_t packed = 0, raw = 0;
  uint64_t rcp = 0;
  bool esc = false;
  auto load_chunk = [&](int64_t c, uint32_t &packed_o, uint32_t &raw_o, uint64_t &rcp_o, bool &esc_o) {
    const int64_t i = c * 32 + lane;
    packed_o = 0; raw_o = 0; rcp_o = 0; esc_o = false;
    if (c >= 0 && i < n) {
      const int32_t t = __ldg(idx + i);
      const int32_t sy = __ldg(sym + i);
      const int32_t maxv = __ldg(cdf_sizes + t) - 2;
      int32_t slot;
      classify(sy, __ldg(offsets + t), maxv, slot, raw_o, esc_o);
      const int32_t *row = cdfs + (int64_t)t * cdf_stride;
      const uint32_t start = (uint32_t)__ldg(row + slot);
      const uint32_t freq = ((uint32_t)__ldg(row + slot + 1) - start) & 0xFFFFu;
      packed_o = (start & 0xFFFFu) | (freq << 16);
      rcp_o = g_rcp_table[freq];
    }
  };
  load_chunk(n_chunks - 1, packed, raw, rcp, esc);
  for (int64_t c = n_chunks - 1; c >= 0; --c) {
    uint32_t packed_n, raw_n;
    uint64_t rcp_n;
    bool esc_n;
    load_chunk(c - 1, packed_n, raw_n, rcp_n, esc_n);
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
      const uint64_t rc = __shfl_sync(0xFFFFFFFFu, rcp, j);
      const uint32_t freq = p >> 16;
      const uint32_t rshift = freq >= 2 ? (uint32_t)(31 - __clz(freq - 1)) : 0u;  // ceil(log2 freq) - 1
      if (enc_put_rcp(x, p & 0xFFFFu, freq, rc, rshift, word)) emit(word);
    }
    packed = packed_n; raw = raw_n; rcp = rcp_n; esc = esc_n;
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
// Decoder CTA = 3 warps per stream:
//   warps 0,1 (producers) turn the known CDF indexes of the NEXT 32 symbols into shared-memory probe windows:
//          for symbol j, lane l gets (start | freq << 16) of slot ws_j + l, where the 32-slot window is centred
//          on the mode of that symbol's table.  None of this depends on the coder state.
//   warp 2 (walker) runs the serial state chain from shared memory only.  Every lane evaluates the complete
//          state update (multiply, renormalise with the speculatively fetched next word) for "its" candidate
//          slot while it tests whether the slot contains the state's low 16 bits; exactly one lane is right and
//          its 63-bit result is broadcast with two warp OR-reductions (redux.sync).  Bit 63 flags the rare
//          events (word consumed / escape symbol); "no lane right" (symbol outside the window) shows as 0.
//          The common path therefore has a single, normally not-taken branch per symbol.
struct DecChunk {
  uint32_t packed[32][32];  // [symbol j][lane]: start | freq << 16 of slot ws_j + lane (0xFFFF = no slot)
  int32_t t[32], size[32], off[32], ws[32];
};

struct WordReader {  // walker-side: lane l caches word[cache_base + l]; `nw` is the next unread word (uniform)
  const uint32_t *base;
  int32_t n_words, rp, cache_base;
  uint32_t cache, nw;
  int lane;
  __device__ __forceinline__ void fill(int32_t from) {
    cache_base = from;
    const int32_t i = from + lane;
    cache = i < n_words ? __ldg(base + i) : 0u;
  }
  __device__ __forceinline__ void init(const uint32_t *b, int32_t n, int l) {
    base = b; n_words = n; lane = l; rp = 0;
    fill(0);
    nw = __shfl_sync(0xFFFFFFFFu, cache, 0);
  }
  __device__ __forceinline__ void consume() {  // warp-uniform
    ++rp;
    if (rp - cache_base >= 32) fill(rp);
    nw = __shfl_sync(0xFFFFFFFFu, cache, rp - cache_base);
  }
  __device__ __forceinline__ uint32_t next() {
    const uint32_t w = nw;
    consume();
    return w;
  }
};

__device__ __forceinline__ int32_t dec_escape(uint64_t &x, WordReader &rd, int32_t maxv) {
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
  int32_t value = (int32_t)(raw >> 1);
  if (raw & 1u) value = -value - 1; else value += maxv;
  return value;
}

constexpr int kDecThreads = 96;

__global__ void __launch_bounds__(kDecThreads)
rans_decode_kernel(const uint8_t *__restrict__ in_bytes, const int64_t *__restrict__ in_starts,
                   const int64_t *__restrict__ in_ends, int n_streams,
                   int64_t n, const int32_t *__restrict__ indexes, const int32_t *__restrict__ cdfs, int cdf_stride,
                   const int32_t *__restrict__ cdf_sizes, const int32_t *__restrict__ offsets,
                   int32_t *__restrict__ out_symbols, const int64_t *__restrict__ seg_start,
                   const int32_t *__restrict__ seg_count) {
  __shared__ DecChunk buf[2];
  __shared__ int32_t s_val[64];  // [0,32): decoded slot per symbol of the chunk; [32,64): dummy slots
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;  // 0,1 = producers (symbols 0-15 / 16-31 of a chunk), 2 = walker
  const int s = blockIdx.x;
  const int64_t first = seg_start ? seg_start[s] : (int64_t)s * n;
  if (seg_count) n = seg_count[s];
  if (n <= 0) return;  // empty segment (uniform for the CTA)
  const int32_t *idx = indexes + first;
  int32_t *out = out_symbols + first;
  const int64_t n_chunks = (n + 31) / 32;

  WordReader rd;
  uint64_t x = 0;
  if (role == 2) {
    rd.init(reinterpret_cast<const uint32_t *>(in_bytes + in_starts[s]), (int32_t)((in_ends[s] - in_starts[s]) / 4), lane);
    x = (uint64_t)rd.next();
    x |= (uint64_t)rd.next() << 32;
  }

  // producers prefetch the index / table-meta loads one chunk ahead of the window build
  int32_t t_n = 0, size_n = 2, off_n = 0;
  auto load_meta = [&](int64_t c) {
    const int64_t i = c * 32 + lane;
    t_n = 0; size_n = 2; off_n = 0;
    if (c < n_chunks && i < n) {
      t_n = __ldg(idx + i);
      size_n = __ldg(cdf_sizes + t_n);
      off_n = __ldg(offsets + t_n);
    }
  };
  if (role < 2) load_meta(0);

  for (int64_t c = 0; c <= n_chunks; ++c) {
    if (role < 2) {
      if (c < n_chunks) {
        DecChunk &d = buf[c & 1];
        const int32_t t_l = t_n, size_l = size_n, off_l = off_n;
        load_meta(c + 1);
        const int32_t ws_l = max(0, min(-off_l - 15, size_l - 32));
        if (role == 0) { d.t[lane] = t_l; d.size[lane] = size_l; d.off[lane] = off_l; d.ws[lane] = ws_l; }
#pragma unroll 8
        for (int jj = 0; jj < 16; ++jj) {
          const int j = role * 16 + jj;
          const int32_t t = __shfl_sync(0xFFFFFFFFu, t_l, j);
          const int32_t size = __shfl_sync(0xFFFFFFFFu, size_l, j);
          const int32_t ws = __shfl_sync(0xFFFFFFFFu, ws_l, j);
          const int32_t *row = cdfs + (int64_t)t * cdf_stride;
          const int32_t pos = ws + lane;
          uint32_t pk = 0xFFFFu;  // start 0xFFFF, freq 0: can never contain a 16-bit cumulative frequency
          if (pos + 1 < size) {
            const uint32_t v = (uint32_t)__ldg(row + pos), v2 = (uint32_t)__ldg(row + pos + 1);
            pk = (v & 0xFFFFu) | ((v2 - v) << 16);
          }
          d.packed[j][lane] = pk;
        }
      }
    } else if (c >= 1) {
      const int64_t cc = c - 1;
      const DecChunk &d = buf[cc & 1];
      const int count = (int)min((int64_t)32, n - cc * 32);
      const int32_t off_l = d.off[lane];
      const int32_t base_l = d.ws[lane] + off_l;  // value = (winning lane) + ws + offset for the lane's own symbol
      uint32_t esc_mask = 0;
      uint32_t pk = d.packed[0][lane];
#pragma unroll 4
      for (int j = 0; j < count; ++j) {
        const uint32_t pk_next = d.packed[min(j + 1, 31)][lane];
        const uint32_t cf = (uint32_t)(x & 0xFFFFu);
        const uint32_t st_l = pk & 0xFFFFu, fr_l = pk >> 16;
        const uint32_t rel = cf - st_l;
        const bool valid = rel < fr_l;
        // candidate successor state of this lane's slot (Rans64DecAdvance + refill with the speculative word)
        uint64_t xc = (uint64_t)fr_l * (x >> kPrecision) + rel;
        const bool need_l = xc < kRansLower;
        if (need_l) xc = (xc << 32) | rd.nw;
        const bool rare_l = need_l || (st_l + fr_l) == 65536u;
        const uint32_t hi_l = (uint32_t)(xc >> 32) | (rare_l ? 0x80000000u : 0u);
        const uint32_t vmask = 0u - (uint32_t)valid;     // branch-free select: all ones for the one right lane
        s_val[valid ? j : 32 + lane] = lane;             // losers write to a private dummy slot (no divergence)
        const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)xc & vmask);
        const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, hi_l & vmask);
        if (__builtin_expect((int32_t)hi >= 0 && (hi | lo) != 0u, 1)) {
          x = ((uint64_t)hi << 32) | lo;  // common case: found in the window, no word consumed, not an escape
        } else {
          bool esc;
          if ((hi | lo) != 0u) {
            x = ((uint64_t)(hi & 0x7FFFFFFFu) << 32) | lo;
            const uint32_t flags = __reduce_or_sync(0xFFFFFFFFu, valid ? ((need_l ? 1u : 0u) |
                                                                           ((st_l + fr_l) == 65536u ? 2u : 0u)) : 0u);
            if (flags & 1u) rd.consume();
            esc = (flags & 2u) != 0u;
          } else {
            // symbol outside the 32-slot window: uniform binary search for the last CDF entry <= cf
            const int32_t size = d.size[j], ws = d.ws[j];
            const int32_t *row = cdfs + (int64_t)d.t[j] * cdf_stride;
            const bool below = ws > 0 && (uint32_t)__ldg(row + ws) > cf;
            int32_t lo_i = below ? 0 : ws;         // cdf[lo] <= cf
            int32_t hi_i = below ? ws : size - 1;  // cdf[hi] > cf
            while (hi_i - lo_i > 1) {
              const int32_t mid = (lo_i + hi_i) >> 1;
              if ((uint32_t)__ldg(row + mid) <= cf) lo_i = mid; else hi_i = mid;
            }
            const uint32_t start = (uint32_t)__ldg(row + lo_i);
            const uint32_t freq = (uint32_t)__ldg(row + lo_i + 1) - start;
            if (dec_advance(x, start, freq)) x = (x << 32) | rd.next();
            esc = (start + freq) == 65536u;
            if (lane == 0) s_val[j] = lo_i - ws;
          }
          if (esc) {  // the last real slot is the escape slot
            const int32_t v = dec_escape(x, rd, d.size[j] - 2);
            __syncwarp();
            if (lane == 0) s_val[j] = v;
            esc_mask |= 1u << j;
          }
        }
        pk = pk_next;
      }
      __syncwarp();
      const int64_t i = cc * 32 + lane;
      if (i < n) out[i] = s_val[lane] + (((esc_mask >> lane) & 1u) ? off_l : base_l);
      __syncwarp();
    }
    __syncthreads();
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
  {
    // one-time fill of the reciprocal table (per device); ordered before the encode on the same stream
    static std::atomic<uint64_t> built_mask{0};
    int dev = 0;
    PCODEC_CHECK_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(built_mask.load() & bit)) {
      rans_build_rcp_kernel<<<256, 256, 0, st>>>();
      PCODEC_COUNT_LAUNCH();
      PCODEC_CHECK_CUDA(cudaStreamSynchronize(st));
      built_mask.fetch_or(bit);
    }
  }
  PCODEC_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  const int blocks = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
  rans_encode_kernel<<<blocks, 32 * kWarpsPerBlock, 0, st>>>(symbols, indexes, n_streams, n_per_stream, cdfs,
                                                              cdf_stride, cdf_sizes, offsets, scratch, scratch_words,
                                                              n_words, status, nullptr, nullptr);
  PCODEC_COUNT_LAUNCH();
  rans_offsets_kernel<<<1, 1024, 0, st>>>(n_words, n_streams, out_offsets, out_cap, status);
  PCODEC_COUNT_LAUNCH();
  rans_compact_kernel<<<dim3(n_streams, 4), 256, 0, st>>>(scratch, scratch_words, n_words, out_offsets, n_streams,
                                                          out_bytes, out_cap);
  PCODEC_COUNT_LAUNCH();
  PCODEC_RETURN_STATUS();
}

extern "C" int pcodec_rans_decode_batch(const uint8_t *in_bytes, const int64_t *in_offsets, int n_streams,
                                        int64_t n_per_stream, const int32_t *indexes, const int32_t *cdfs,
                                        int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int n_tables,
                                        int32_t *out_symbols, void *stream) {
  (void)n_tables;
  if (n_streams <= 0 || n_per_stream < 0 || !in_bytes || !in_offsets) return PCODEC_ERR_BAD_ARG;
  if (n_per_stream == 0) return PCODEC_OK;
  if (!indexes || !out_symbols) return PCODEC_ERR_BAD_ARG;
  rans_decode_kernel<<<n_streams, kDecThreads, 0, as_stream(stream)>>>(
      in_bytes, in_offsets, in_offsets + 1, n_streams, n_per_stream, indexes, cdfs, cdf_stride, cdf_sizes, offsets,
      out_symbols, nullptr, nullptr);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_rans_decode_ranges(const uint8_t *in_bytes, const int64_t *starts, const int64_t *ends,
                                         int n_streams, int64_t n_per_stream, const int32_t *indexes,
                                         const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                         const int32_t *offsets, int n_tables, int32_t *out_symbols, void *stream) {
  (void)n_tables;
  if (n_streams <= 0 || n_per_stream < 0 || !in_bytes || !starts || !ends) return PCODEC_ERR_BAD_ARG;
  if (n_per_stream == 0) return PCODEC_OK;
  if (!indexes || !out_symbols) return PCODEC_ERR_BAD_ARG;
  rans_decode_kernel<<<n_streams, kDecThreads, 0, as_stream(stream)>>>(
      in_bytes, starts, ends, n_streams, n_per_stream, indexes, cdfs, cdf_stride, cdf_sizes, offsets, out_symbols,
      nullptr, nullptr);
  PCODEC_RETURN_LAUNCH();
}

// Variable-length streams ("segments") of flat symbol / index arrays: the progressive container codes, per (layer,
// slice, image), only the latent elements that ENTER at that layer.
extern "C" int pcodec_rans_encode_segments(const int32_t *symbols, const int32_t *indexes, const int64_t *seg_start,
                                           const int32_t *seg_count, int n_streams, const int32_t *cdfs, int cdf_stride,
                                           const int32_t *cdf_sizes, const int32_t *offsets, int n_tables,
                                           uint32_t *scratch, int64_t scratch_words, int32_t *n_words,
                                           uint8_t *out_bytes, int64_t out_cap, int64_t *out_offsets, int32_t *status,
                                           void *stream) {
  (void)n_tables;
  if (n_streams <= 0 || scratch_words < 2 || !scratch || !n_words || !out_bytes || !out_offsets || !status || !symbols ||
      !indexes || !seg_start || !seg_count)
    return PCODEC_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  {
    static std::atomic<uint64_t> built_mask{0};
    int dev = 0;
    PCODEC_CHECK_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(built_mask.load() & bit)) {
      rans_build_rcp_kernel<<<256, 256, 0, st>>>();
      PCODEC_COUNT_LAUNCH();
      PCODEC_CHECK_CUDA(cudaStreamSynchronize(st));
      built_mask.fetch_or(bit);
    }
  }
  PCODEC_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  rans_encode_kernel<<<n_streams, 32 * kWarpsPerBlock, 0, st>>>(symbols, indexes, n_streams, 0, cdfs, cdf_stride,
                                                                 cdf_sizes, offsets, scratch, scratch_words, n_words,
                                                                 status, seg_start, seg_count);
  PCODEC_COUNT_LAUNCH();
  rans_offsets_kernel<<<1, 1024, 0, st>>>(n_words, n_streams, out_offsets, out_cap, status);
  PCODEC_COUNT_LAUNCH();
  rans_compact_kernel<<<dim3(n_streams, 4), 256, 0, st>>>(scratch, scratch_words, n_words, out_offsets, n_streams,
                                                          out_bytes, out_cap);
  PCODEC_COUNT_LAUNCH();
  PCODEC_RETURN_STATUS();
}

extern "C" int pcodec_rans_decode_segments(const uint8_t *in_bytes, const int64_t *starts, const int64_t *ends,
                                           int n_streams, const int64_t *seg_start, const int32_t *seg_count,
                                           const int32_t *indexes, const int32_t *cdfs, int cdf_stride,
                                           const int32_t *cdf_sizes, const int32_t *offsets, int n_tables,
                                           int32_t *out_symbols, void *stream) {
  (void)n_tables;
  if (n_streams <= 0 || !in_bytes || !starts || !ends || !seg_start || !seg_count || !indexes || !out_symbols)
    return PCODEC_ERR_BAD_ARG;
  rans_decode_kernel<<<n_streams, kDecThreads, 0, as_stream(stream)>>>(
      in_bytes, starts, ends, n_streams, 0, indexes, cdfs, cdf_stride, cdf_sizes, offsets, out_symbols, seg_start,
      seg_count);
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
    // alternate between the division and the reciprocal form: both must yield the oracle's bytes
    const bool e = (i & 1) ? enc_put(x, start & 0xFFFFu, freq, word)
                           : enc_put_rcp(x, start & 0xFFFFu, freq, rcp_for(freq), freq >= 2 ? rcp_shift_for(freq) : 0u, word);
    if (e) { if (wpos <= 0) return -1; words[--wpos] = word; }
  }
  if (wpos < 2) return -1;
  words[--wpos] = (uint32_t)(x >> 32);
  words[--wpos] = (uint32_t)x;
  return cap_words - wpos;  // words used, stored at the END of `words`
}
