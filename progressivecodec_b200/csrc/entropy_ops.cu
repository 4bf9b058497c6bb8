// Variance-aware masking, quantisation, CDF-index lookup and likelihoods (bandwidth-class kernels).
//
// Reference semantics (paths under /root/reference/src/compress/):
//   layers/masking.py:205-223            per-image torch.quantile of sigma, mask = sigma >= threshold
//   entropy_models/entropy_models.py:126-165   quantize ("symbols" / "dequantize"), dequantize
//   entropy_models/entropy_models.py:661-666   build_indexes over the 64-level scale table
//   entropy_models/entropy_models.py:626-659   GaussianConditional likelihood
//   entropy_models/entropy_models.py:400-433   EntropyBottleneck factorised density
//
// Data movement: activations are NHWC (32 channels of a pixel = one 128-byte line), while the entropy
// coder consumes planes in NCHW order.  Each CTA therefore handles a [32 pixels x 32 channels] tile:
// coalesced 128 B reads along channels, a padded shared-memory transpose, coalesced 128 B writes along
// pixels.  Every element is read once and every output written once (24 B/element for the encoder step).
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kTile = 32;

__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// ------------------------------------------------------------------------------------------------
// torch.quantile threshold per image: 4-pass 8-bit radix select of the two order statistics around
// rank = q*(n-1), then ATen's lerp.  One CTA per image.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
quantile_threshold_kernel(const float *__restrict__ scale, int64_t hw, int channels, int ps, float q,
                          float *__restrict__ thr, int cache_keys) {
  // cache_keys: the n sort keys of the image fit the dynamic shared memory (196 KB for a 32x48x32 slice), so the four
  // radix passes and the successor search read them from there instead of re-reading sigma from L2 five times
  extern __shared__ uint32_t s_keys[];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_digit, s_kk, s_below, s_equal, s_min;
  const int tid = threadIdx.x, lane = tid & 31;
  const float *img = scale + (int64_t)blockIdx.x * hw * ps;
  const int64_t n = hw * channels;

  const float rank = __fmul_rn(q, (float)(n - 1));
  const float lo_f = floorf(rank);
  const float w = __fsub_rn(rank, lo_f);
  int64_t k = (int64_t)lo_f;
  if (k < 0) k = 0;
  if (k > n - 1) k = n - 1;

  uint32_t prefix = 0, pmask = 0;
  if (tid == 0) { s_kk = (uint32_t)k; s_below = 0; }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int64_t e0 = 0; e0 < n; e0 += 1024) {
      const int64_t e = e0 + tid;
      int bin = -1;
      if (e < n) {
        uint32_t key;
        if (cache_keys && pass > 0) {
          key = s_keys[e];
        } else {
          const uint32_t ee = (uint32_t)e, px = ee / (uint32_t)channels;
          key = float_key(__ldg(img + (int64_t)px * ps + (ee - px * (uint32_t)channels)));
          if (cache_keys) s_keys[e] = key;
        }
        if ((key & pmask) == prefix) bin = (int)((key >> shift) & 255u);
      }
      // warp-aggregated histogram update (sigma keys cluster in a handful of top-digit bins)
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, bin);
      if (bin >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t kk = s_kk, cum = 0;
      int d = 0;
      for (; d < 255; ++d) {
        if (cum + hist[d] > kk) break;
        cum += hist[d];
      }
      s_digit = (uint32_t)d;
      s_kk = kk - cum;
      s_below += cum;
      s_equal = hist[d];
    }
    __syncthreads();
    prefix |= s_digit << shift;
    pmask |= 0xFFu << shift;
    __syncthreads();
  }
  const uint32_t key_lo = prefix;
  const int64_t below = s_below, equal = s_equal;
  uint32_t key_hi = key_lo;
  if (k + 1 > below + equal - 1 && k + 1 <= n - 1) {  // next order statistic is the smallest key > key_lo
    if (tid == 0) s_min = 0xFFFFFFFFu;
    __syncthreads();
    uint32_t best = 0xFFFFFFFFu;
    for (int64_t e = tid; e < n; e += 1024) {
      uint32_t key;
      if (cache_keys) {
        key = s_keys[e];
      } else {
        const uint32_t ee = (uint32_t)e, px = ee / (uint32_t)channels;
        key = float_key(__ldg(img + (int64_t)px * ps + (ee - px * (uint32_t)channels)));
      }
      if (key > key_lo && key < best) best = key;
    }
    for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
    if (lane == 0) atomicMin(&s_min, best);
    __syncthreads();
    key_hi = s_min;
  }
  if (tid == 0) {
    const float a = key_float(key_lo), b = key_float(key_hi);
    // ATen lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w); no FMA contraction
    const float d = __fsub_rn(b, a);
    float r;
    if (w < 0.5f) r = __fadd_rn(a, __fmul_rn(w, d));
    else r = __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, w)));
    thr[blockIdx.x] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// helpers for the tile kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int scale_index(const float *s_table, int n_thresholds, float s) {
  // number of thresholds t_j (j < n_thresholds) with t_j < s  ==  63 - sum_j [s <= t_j]
  int lo = 0, hi = n_thresholds;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_table[mid] < s) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float std_cumulative(float x) {
  return 0.5f * erfcf(-0.70710678118654752440f * x);
}

struct QuantArgs {
  const float *y; int y_ps;
  const float *y_sub; int y_sub_ps;
  const float *mu; int mu_ps;
  const float *scale; int scale_ps;
  int64_t hw; int channels; int mask_mode;
  const float *thr; const float *table; int n_levels; float bound;
  int32_t *symbols; int32_t *indexes; float *mask_out; float *lik;
  float *y_hat; int y_hat_ps;
  const float *mask_src; int mask_src_ps;  // cust_map: the mask is mask_src >= thr instead of scale >= thr
};

// grid: (pixel tiles, channel groups of 32, batch); block (32, 8)
__global__ void __launch_bounds__(256) slice_quantize_kernel(QuantArgs a) {
  extern __shared__ float s_table[];
  __shared__ int32_t t_sym[kTile][kTile + 1];
  __shared__ int32_t t_idx[kTile][kTile + 1];
  __shared__ float t_msk[kTile][kTile + 1];
  __shared__ float t_lik[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty * 32 + tx; i < a.n_levels; i += 256) s_table[i] = a.table[i];
  __syncthreads();
  const int b = blockIdx.z;
  const int c = blockIdx.y * kTile + tx;
  const int64_t p0 = (int64_t)blockIdx.x * kTile;
  const float thr = (a.mask_mode == PCODEC_MASK_THRESHOLD) ? a.thr[b] : 0.0f;
  for (int pp = ty; pp < kTile; pp += 8) {
    const int64_t p = p0 + pp;
    int32_t sym = 0, idx = 0;
    float m = 0.f, lk = 1.f;
    if (p < a.hw && c < a.channels) {
      const int64_t pix = (int64_t)b * a.hw + p;
      float s = a.scale ? a.scale[pix * a.scale_ps + c] : 0.f;
      const float mu = a.mu ? a.mu[pix * a.mu_ps + c] : 0.f;
      const float ms = a.mask_src ? a.mask_src[pix * a.mask_src_ps + c] : s;
      m = a.mask_mode == PCODEC_MASK_ONES ? 1.f : (a.mask_mode == PCODEC_MASK_ZEROS ? 0.f : (ms >= thr ? 1.f : 0.f));
      float r = 0.f;
      if (a.y) {
        float v = a.y[pix * a.y_ps + c];
        if (a.y_sub) v = __fsub_rn(v, a.y_sub[pix * a.y_sub_ps + c]);
        v = __fsub_rn(v, mu);
        r = rintf(__fmul_rn(v, m));
        sym = (int32_t)r;
        if (a.y_hat) a.y_hat[pix * a.y_hat_ps + c] = __fadd_rn(r, mu);
      }
      const float sb = fmaxf(__fmul_rn(s, m), a.bound);
      idx = scale_index(s_table, a.n_levels - 1, sb);
      if (a.lik) {
        const float av = fabsf(r);
        const float up = std_cumulative(__fdiv_rn(0.5f - av, sb));
        const float lw = std_cumulative(__fdiv_rn(-0.5f - av, sb));
        lk = fmaxf(up - lw, 1e-9f);
      }
    }
    t_sym[tx][pp] = sym;
    t_idx[tx][pp] = idx;
    t_msk[tx][pp] = m;
    t_lik[tx][pp] = lk;
  }
  __syncthreads();
  // NCHW writes: tx = pixel within tile, ty strides over channels
  const int64_t p = p0 + tx;
  if (p < a.hw) {
    for (int cc = ty; cc < kTile; cc += 8) {
      const int ch = blockIdx.y * kTile + cc;
      if (ch >= a.channels) break;
      const int64_t o = ((int64_t)b * a.channels + ch) * a.hw + p;
      if (a.symbols) a.symbols[o] = t_sym[cc][tx];
      if (a.indexes) a.indexes[o] = t_idx[cc][tx];
      if (a.mask_out) a.mask_out[o] = t_msk[cc][tx];
      if (a.lik) a.lik[o] = t_lik[cc][tx];
    }
  }
}

// y_hat(NHWC) = float(sym NCHW) + mu(NHWC): read NCHW along pixels, transpose, write NHWC along channels.
__global__ void __launch_bounds__(256)
slice_dequantize_kernel(const int32_t *__restrict__ symbols, const float *__restrict__ mu, int mu_ps,
                        const float *__restrict__ per_channel_add, int64_t hw, int channels, float *__restrict__ y_hat,
                        int y_hat_ps) {
  __shared__ float t[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * kTile;
  {
    const int64_t p = p0 + tx;
    for (int cc = ty; cc < kTile; cc += 8) {
      const int ch = blockIdx.y * kTile + cc;
      float v = 0.f;
      if (p < hw && ch < channels) v = (float)symbols[((int64_t)b * channels + ch) * hw + p];
      t[cc][tx] = v;
    }
  }
  __syncthreads();
  const int c = blockIdx.y * kTile + tx;
  if (c < channels) {
    for (int pp = ty; pp < kTile; pp += 8) {
      const int64_t p = p0 + pp;
      if (p >= hw) break;
      const int64_t pix = (int64_t)b * hw + p;
      float add = mu ? mu[pix * mu_ps + c] : 0.f;
      if (per_channel_add) add = per_channel_add[c];
      y_hat[pix * y_hat_ps + c] = __fadd_rn(t[tx][pp], add);
    }
  }
}

// z path: sym = rint(z - median[c]); index = c; z_hat = sym + median[c]
__global__ void __launch_bounds__(256)
bottleneck_quantize_kernel(const float *__restrict__ z, int z_ps, const float *__restrict__ medians, int64_t hw,
                           int channels, int32_t *__restrict__ symbols, int32_t *__restrict__ indexes,
                           float *__restrict__ z_hat, int z_hat_ps) {
  __shared__ int32_t t_sym[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
  const int c = blockIdx.y * kTile + tx;
  const int64_t p0 = (int64_t)blockIdx.x * kTile;
  for (int pp = ty; pp < kTile; pp += 8) {
    const int64_t p = p0 + pp;
    int32_t sym = 0;
    if (p < hw && c < channels) {
      const int64_t pix = (int64_t)b * hw + p;
      const float med = medians[c];
      const float r = rintf(__fsub_rn(z[pix * z_ps + c], med));
      sym = (int32_t)r;
      if (z_hat) z_hat[pix * z_hat_ps + c] = __fadd_rn(r, med);
    }
    t_sym[tx][pp] = sym;
  }
  __syncthreads();
  const int64_t p = p0 + tx;
  if (p < hw) {
    for (int cc = ty; cc < kTile; cc += 8) {
      const int ch = blockIdx.y * kTile + cc;
      if (ch >= channels) break;
      const int64_t o = ((int64_t)b * channels + ch) * hw + p;
      if (symbols) symbols[o] = t_sym[cc][tx];
      if (indexes) indexes[o] = ch;
    }
  }
}

__global__ void bottleneck_indexes_kernel(int64_t hw, int channels, int64_t total, int32_t *__restrict__ indexes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) indexes[i] = (int32_t)((i / hw) % channels);
}

__device__ __forceinline__ float logits_cumulative(const float *p, float v) {
  // p: softplus(m0)[3] m1[9] m2[9] m3[9] m4[3] | b0[3] b1[3] b2[3] b3[3] b4[1] | tanh(f0..f3)[3 each]
  const float *m0 = p, *m1 = p + 3, *m2 = p + 12, *m3 = p + 21, *m4 = p + 30;
  const float *b0 = p + 33, *b1 = p + 36, *b2 = p + 39, *b3 = p + 42, *b4 = p + 45;
  const float *f0 = p + 46, *f1 = p + 49, *f2 = p + 52, *f3 = p + 55;
  float l[3], t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    l[r] = __fadd_rn(__fmul_rn(m0[r], v), b0[r]);
    l[r] = __fadd_rn(l[r], __fmul_rn(f0[r], tanhf(l[r])));
  }
  const float *ms[3] = {m1, m2, m3};
  const float *bs[3] = {b1, b2, b3};
  const float *fs[3] = {f1, f2, f3};
#pragma unroll
  for (int layer = 0; layer < 3; ++layer) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = __fmul_rn(ms[layer][r * 3 + 0], l[0]);
      acc = __fadd_rn(acc, __fmul_rn(ms[layer][r * 3 + 1], l[1]));
      acc = __fadd_rn(acc, __fmul_rn(ms[layer][r * 3 + 2], l[2]));
      acc = __fadd_rn(acc, bs[layer][r]);
      t[r] = __fadd_rn(acc, __fmul_rn(fs[layer][r], tanhf(acc)));
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) l[r] = t[r];
  }
  float acc = __fmul_rn(m4[0], l[0]);
  acc = __fadd_rn(acc, __fmul_rn(m4[1], l[1]));
  acc = __fadd_rn(acc, __fmul_rn(m4[2], l[2]));
  return __fadd_rn(acc, b4[0]);
}

__global__ void __launch_bounds__(256)
bottleneck_likelihood_kernel(const float *__restrict__ z_hat, int z_ps, const float *__restrict__ params, int64_t hw,
                             int channels, float *__restrict__ lik) {
  __shared__ float t[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
  const int c = blockIdx.y * kTile + tx;
  const int64_t p0 = (int64_t)blockIdx.x * kTile;
  for (int pp = ty; pp < kTile; pp += 8) {
    const int64_t p = p0 + pp;
    float lk = 1.f;
    if (p < hw && c < channels) {
      const float v = z_hat[((int64_t)b * hw + p) * z_ps + c];
      const float *prm = params + (int64_t)c * 58;
      const float lower = logits_cumulative(prm, v - 0.5f);
      const float upper = logits_cumulative(prm, v + 0.5f);
      const float sum = lower + upper;
      const float sign = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
      lk = fabsf(sigmoid_f(sign * upper) - sigmoid_f(sign * lower));
      lk = fmaxf(lk, 1e-9f);
    }
    t[tx][pp] = lk;
  }
  __syncthreads();
  const int64_t p = p0 + tx;
  if (p < hw) {
    for (int cc = ty; cc < kTile; cc += 8) {
      const int ch = blockIdx.y * kTile + cc;
      if (ch >= channels) break;
      lik[((int64_t)b * channels + ch) * hw + p] = t[cc][tx];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float *__restrict__ src, float *__restrict__ dst, int channels, int64_t hw, int dst_ps,
                    int c_pad) {
  __shared__ float t[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * kTile;
  const int64_t p = p0 + tx;
  for (int cc = ty; cc < kTile; cc += 8) {
    const int ch = blockIdx.y * kTile + cc;
    t[cc][tx] = (p < hw && ch < channels) ? src[((int64_t)b * channels + ch) * hw + p] : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.y * kTile + tx;
  if (c < c_pad) {
    for (int pp = ty; pp < kTile; pp += 8) {
      const int64_t q = p0 + pp;
      if (q >= hw) break;
      dst[((int64_t)b * hw + q) * dst_ps + c] = t[tx][pp];
    }
  }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float *__restrict__ src, int src_ps, float *__restrict__ dst, int channels, int64_t hw) {
  __shared__ float t[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * kTile;
  const int c = blockIdx.y * kTile + tx;
  for (int pp = ty; pp < kTile; pp += 8) {
    const int64_t q = p0 + pp;
    t[tx][pp] = (q < hw && c < channels) ? src[((int64_t)b * hw + q) * src_ps + c] : 0.f;
  }
  __syncthreads();
  const int64_t p = p0 + tx;
  if (p < hw) {
    for (int cc = ty; cc < kTile; cc += 8) {
      const int ch = blockIdx.y * kTile + cc;
      if (ch >= channels) break;
      dst[((int64_t)b * channels + ch) * hw + p] = t[cc][tx];
    }
  }
}

inline dim3 tile_grid(int64_t hw, int channels, int batch) {
  return dim3((unsigned)ceil_div64(hw, kTile), (unsigned)((channels + kTile - 1) / kTile), (unsigned)batch);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Progressive-layer partition (truncatable container)
// ------------------------------------------------------------------------------------------------
// The variance-aware masks of increasing quality levels are nested (thr_0 >= thr_1 >= ...), so every latent element
// of a progressive slice ENTERS at exactly one level: layer(e) = #{k : sigma_e < thr_k}.  One CTA per image stably
// partitions the slice's elements (in the coder's NCHW order) by layer:
//   gather : raster planes (symbols, indexes) -> layer-major compacted arrays + per-layer counts
//   scatter: decoded compacted symbols of the first `avail` layers -> raster plane (zeros elsewhere)
// Both directions compute (layer, rank) identically, so the decoder recovers the encoder's order from sigma alone.
constexpr int kMaxLayers = 15;  // + 1 bucket for "never enters"

template <bool SCATTER>
__global__ void __launch_bounds__(1024)
layer_partition_kernel(const float *__restrict__ sigma, int sigma_ps, int64_t hw, int channels,
                       const float *__restrict__ thr /* [n_levels][batch] */, int n_levels, int batch,
                       const int32_t *__restrict__ in_a, const int32_t *__restrict__ in_b,
                       int32_t *__restrict__ out_a, int32_t *__restrict__ out_b, int32_t *__restrict__ counts,
                       const int32_t *__restrict__ avail) {
  __shared__ int s_hist[kMaxLayers + 1];
  __shared__ int s_base[kMaxLayers + 1];
  __shared__ int s_warp[32][kMaxLayers + 1];
  __shared__ float s_thr[kMaxLayers];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = hw * channels;
  const float *sg = sigma + (int64_t)b * hw * sigma_ps;
  if (tid <= kMaxLayers) s_hist[tid] = 0;
  if (tid < n_levels) s_thr[tid] = thr[(int64_t)tid * batch + b];
  __syncthreads();
  auto layer_of = [&](int64_t e) {  // element e = (channel c, pixel p) of the NCHW plane
    const int64_t c = e / hw, p = e - c * hw;
    const float v = sg[p * sigma_ps + c];
    int k = 0;
    for (int l = 0; l < n_levels; ++l) k += (v < s_thr[l]) ? 1 : 0;  // NaN thresholds (pr >= 10: "ones") never exclude
    return k;
  };
  // pass 1: histogram
  for (int64_t e = tid; e < n; e += blockDim.x) atomicAdd(&s_hist[layer_of(e)], 1);
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int k = 0; k <= n_levels; ++k) { s_base[k] = run; run += s_hist[k]; }
  }
  __syncthreads();
  if (!SCATTER && counts && tid <= n_levels) counts[(int64_t)b * (kMaxLayers + 1) + tid] = s_hist[tid];
  const int n_avail = SCATTER ? avail[b] : 0;
  const int64_t row = (int64_t)b * n;
  // pass 2: chunks of 1024 consecutive elements; rank inside the chunk by per-layer ballots
  for (int64_t e0 = 0; e0 < n; e0 += blockDim.x) {
    const int64_t e = e0 + tid;
    const bool ok = e < n;
    const int k = ok ? layer_of(e) : -1;
    int my_rank = 0;
    for (int l = 0; l <= n_levels; ++l) {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, k == l);
      if (k == l) my_rank = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) s_warp[warp][l] = __popc(m);
    }
    __syncthreads();
    if (ok) {
      int before = 0;
      for (int w = 0; w < warp; ++w) before += s_warp[w][k];
      const int64_t pos = row + s_base[k] + before + my_rank;
      if (!SCATTER) {
        if (in_a) out_a[pos] = in_a[row + e];
        if (in_b) out_b[pos] = in_b[row + e];
      } else {
        out_a[row + e] = k < n_avail ? in_a[pos] : 0;
      }
    }
    __syncthreads();
    if (tid <= n_levels) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w][tid];
      s_base[tid] += tot;
    }
    __syncthreads();
  }
}

extern "C" int pcodec_quantile_threshold(const float *scale, int batch, int64_t hw, int channels, int pixel_stride,
                                         float q, float *thr, uint32_t *workspace, void *stream) {
  (void)workspace;
  if (!scale || !thr || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  const int64_t n = hw * channels;
  if (n >= (1ll << 31)) return PCODEC_ERR_UNSUPPORTED;
  constexpr int kKeyCacheMax = 200 * 1024;  // bytes of dynamic shared memory for the key cache
  const int cache = n * 4 <= kKeyCacheMax ? 1 : 0;
  {  // opt-in to the large dynamic shared memory once per DEVICE (the attribute applies to the current device only)
    static std::atomic<uint64_t> attr_mask{0};
    uint64_t bit;
    if (pcodec_device_needs(attr_mask, &bit)) {
      PCODEC_CHECK_CUDA(cudaFuncSetAttribute(quantile_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKeyCacheMax));
      attr_mask.fetch_or(bit, std::memory_order_release);
    }
  }
  quantile_threshold_kernel<<<batch, 1024, cache ? (size_t)n * 4 : 0, as_stream(stream)>>>(scale, hw, channels,
                                                                                            pixel_stride, q, thr, cache);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_slice_quantize_cust(const float *y, int y_ps, const float *y_sub, int y_sub_ps, const float *mu,
                                          int mu_ps, const float *scale, int scale_ps, int batch, int64_t hw, int channels,
                                          int mask_mode, const float *thr, const float *scale_table, int n_levels,
                                          float scale_bound, int32_t *symbols, int32_t *indexes, float *mask_out,
                                          float *lik, float *y_hat, int y_hat_ps, const float *mask_src, int mask_src_ps,
                                          void *stream) {
  if (batch <= 0 || hw <= 0 || channels <= 0 || !scale_table || n_levels < 1 || n_levels > 4096)
    return PCODEC_ERR_BAD_ARG;
  if (mask_mode == PCODEC_MASK_THRESHOLD && (!thr || (!scale && !mask_src))) return PCODEC_ERR_BAD_ARG;
  QuantArgs a{y, y_ps, y_sub, y_sub_ps, mu, mu_ps, scale, scale_ps, hw, channels, mask_mode, thr, scale_table,
              n_levels, scale_bound, symbols, indexes, mask_out, lik, y_hat, y_hat_ps, mask_src, mask_src_ps};
  slice_quantize_kernel<<<tile_grid(hw, channels, batch), dim3(32, 8), sizeof(float) * n_levels, as_stream(stream)>>>(a);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_slice_quantize(const float *y, int y_ps, const float *y_sub, int y_sub_ps, const float *mu,
                                     int mu_ps, const float *scale, int scale_ps, int batch, int64_t hw, int channels,
                                     int mask_mode, const float *thr, const float *scale_table, int n_levels,
                                     float scale_bound, int32_t *symbols, int32_t *indexes, float *mask_out, float *lik,
                                     float *y_hat, int y_hat_ps, void *stream) {
  return pcodec_slice_quantize_cust(y, y_ps, y_sub, y_sub_ps, mu, mu_ps, scale, scale_ps, batch, hw, channels, mask_mode, thr,
                                    scale_table, n_levels, scale_bound, symbols, indexes, mask_out, lik, y_hat, y_hat_ps,
                                    nullptr, 0, stream);
}

extern "C" int pcodec_slice_indexes(const float *scale, int scale_ps, int batch, int64_t hw, int channels,
                                    int mask_mode, const float *thr, const float *scale_table, int n_levels,
                                    float scale_bound, int32_t *indexes, void *stream) {
  return pcodec_slice_quantize(nullptr, 0, nullptr, 0, nullptr, 0, scale, scale_ps, batch, hw, channels, mask_mode, thr,
                               scale_table, n_levels, scale_bound, nullptr, indexes, nullptr, nullptr, nullptr, 0,
                               stream);
}

extern "C" int pcodec_slice_dequantize(const int32_t *symbols, const float *mu, int mu_ps, int batch, int64_t hw,
                                       int channels, float *y_hat, int y_hat_ps, void *stream) {
  if (!symbols || !y_hat || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  slice_dequantize_kernel<<<tile_grid(hw, channels, batch), dim3(32, 8), 0, as_stream(stream)>>>(
      symbols, mu, mu_ps, nullptr, hw, channels, y_hat, y_hat_ps);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_bottleneck_quantize(const float *z, int z_ps, const float *medians, int batch, int64_t hw,
                                          int channels, int32_t *symbols, int32_t *indexes, float *z_hat, int z_hat_ps,
                                          void *stream) {
  if (!z || !medians || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  bottleneck_quantize_kernel<<<tile_grid(hw, channels, batch), dim3(32, 8), 0, as_stream(stream)>>>(
      z, z_ps, medians, hw, channels, symbols, indexes, z_hat, z_hat_ps);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_bottleneck_dequantize(const int32_t *symbols, const float *medians, int batch, int64_t hw,
                                            int channels, float *z_hat, int z_hat_ps, void *stream) {
  if (!symbols || !medians || !z_hat || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  slice_dequantize_kernel<<<tile_grid(hw, channels, batch), dim3(32, 8), 0, as_stream(stream)>>>(
      symbols, nullptr, 0, medians, hw, channels, z_hat, z_hat_ps);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_bottleneck_indexes(int batch, int64_t hw, int channels, int32_t *indexes, void *stream) {
  if (!indexes || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  const int64_t total = (int64_t)batch * channels * hw;
  bottleneck_indexes_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(hw, channels, total,
                                                                                            indexes);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_bottleneck_likelihood(const float *z_hat, int z_ps, const float *params, int batch, int64_t hw,
                                            int channels, float *lik, void *stream) {
  if (!z_hat || !params || !lik || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  bottleneck_likelihood_kernel<<<tile_grid(hw, channels, batch), dim3(32, 8), 0, as_stream(stream)>>>(
      z_hat, z_ps, params, hw, channels, lik);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_nchw_to_nhwc(const float *src, float *dst, int batch, int channels, int64_t hw, int dst_ps,
                                   int c_pad, void *stream) {
  if (!src || !dst || batch <= 0 || hw <= 0 || channels <= 0 || c_pad < channels || dst_ps < c_pad)
    return PCODEC_ERR_BAD_ARG;
  nchw_to_nhwc_kernel<<<tile_grid(hw, c_pad, batch), dim3(32, 8), 0, as_stream(stream)>>>(src, dst, channels, hw,
                                                                                           dst_ps, c_pad);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_nhwc_to_nchw(const float *src, int src_ps, float *dst, int batch, int channels, int64_t hw,
                                   void *stream) {
  if (!src || !dst || batch <= 0 || hw <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  nhwc_to_nchw_kernel<<<tile_grid(hw, channels, batch), dim3(32, 8), 0, as_stream(stream)>>>(src, src_ps, dst,
                                                                                             channels, hw);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_layer_partition(const float *sigma, int sigma_ps, int batch, int64_t hw, int channels,
                                      const float *thresholds, int n_levels, const int32_t *in_a, const int32_t *in_b,
                                      int32_t *out_a, int32_t *out_b, int32_t *counts, const int32_t *avail,
                                      int scatter, void *stream) {
  if (!sigma || !thresholds || batch <= 0 || hw <= 0 || channels <= 0 || n_levels < 1 || n_levels > kMaxLayers)
    return PCODEC_ERR_BAD_ARG;
  if (scatter) {
    if (!in_a || !out_a || !avail) return PCODEC_ERR_BAD_ARG;
    layer_partition_kernel<true><<<batch, 1024, 0, as_stream(stream)>>>(sigma, sigma_ps, hw, channels, thresholds,
                                                                          n_levels, batch, in_a, nullptr, out_a, nullptr,
                                                                          nullptr, avail);
  } else {
    if ((in_a && !out_a) || (in_b && !out_b) || !counts) return PCODEC_ERR_BAD_ARG;
    layer_partition_kernel<false><<<batch, 1024, 0, as_stream(stream)>>>(sigma, sigma_ps, hw, channels, thresholds,
                                                                           n_levels, batch, in_a, in_b, out_a, out_b,
                                                                           counts, nullptr);
  }
  PCODEC_RETURN_LAUNCH();
}

// out = ret * (star - bar) + identity  (REM wrapper; see include/pcodec_b200.h)
__global__ void __launch_bounds__(256)
masked_residual_kernel(const float *__restrict__ ret, int ret_ps, const float *__restrict__ identity, int id_ps,
                       const float *__restrict__ sigma, int sigma_ps, int64_t hw, int channels, int mode_star,
                       const float *__restrict__ thr_star, int mode_bar, const float *__restrict__ thr_bar,
                       float *__restrict__ out, int out_ps, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % channels);
  const int64_t pix = e / channels;
  const int b = (int)(pix / hw);
  const float s = sigma[pix * sigma_ps + (c & 31)];
  const float star = mode_star == PCODEC_MASK_ONES ? 1.f : (mode_star == PCODEC_MASK_ZEROS ? 0.f : (s >= thr_star[b] ? 1.f : 0.f));
  const float bar = mode_bar == PCODEC_MASK_ONES ? 1.f : (mode_bar == PCODEC_MASK_ZEROS ? 0.f : (s >= thr_bar[b] ? 1.f : 0.f));
  const float att = rintf(__fsub_rn(star, bar));  // apply_noise(mask, training=False) = torch.round
  out[pix * out_ps + c] = __fadd_rn(__fmul_rn(ret[pix * ret_ps + c], att), identity[pix * id_ps + c]);
}

extern "C" int pcodec_masked_residual(const float *ret, int ret_ps, const float *identity, int id_ps, const float *sigma,
                                      int sigma_ps, int batch, int64_t hw, int channels, int mode_star,
                                      const float *thr_star, int mode_bar, const float *thr_bar, float *out, int out_ps,
                                      void *stream) {
  if (!ret || !identity || !sigma || !out || batch <= 0 || hw <= 0 || (channels != 32 && channels != 64))
    return PCODEC_ERR_BAD_ARG;
  if ((mode_star == PCODEC_MASK_THRESHOLD && !thr_star) || (mode_bar == PCODEC_MASK_THRESHOLD && !thr_bar))
    return PCODEC_ERR_BAD_ARG;
  const int64_t total = (int64_t)batch * hw * channels;
  masked_residual_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(
      ret, ret_ps, identity, id_ps, sigma, sigma_ps, hw, channels, mode_star, thr_star, mode_bar, thr_bar, out, out_ps,
      total);
  PCODEC_RETURN_LAUNCH();
}
