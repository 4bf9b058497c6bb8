// Shifted-window multi-head self-attention core (reference: layers/win_attention.py:84-115, 153-207).
//
// The qkv and proj Linear layers are 1x1 convolutions on the NHWC tensor (pcodec_conv_taps); this kernel does
// everything between them without materialising torch.roll / window_partition / window_reverse copies: a
// (window, head) CTA gathers its T = ws*ws tokens straight from the rolled pixel positions, adds the
// relative-position bias and the SW-MSA region mask (-100), soft-maxes and writes P.V back to the un-rolled
// pixel.  T is 64 (ws 8) or 16 (ws 4); head_dim 24 / 40 / 80.
#include "common.cuh"

namespace {

template <int T>
__global__ void __launch_bounds__(T)
window_attention_kernel(const float *__restrict__ qkv, int qkv_ps, float *__restrict__ out, int out_ps,
                        const float *__restrict__ rel_bias, int H, int W, int C, int heads, int ws, int shift) {
  extern __shared__ float smem[];
  const int hd = C / heads;
  const int ldk = hd + 1;
  float *sq = smem;             // [T][ldk]
  float *sk = sq + T * ldk;     // [T][ldk]
  float *sv = sk + T * ldk;     // [T][hd]
  __shared__ int s_region[T];

  const int nww = W / ws, nwh = H / ws;
  const int win = blockIdx.x % (nwh * nww);
  const int b = blockIdx.x / (nwh * nww);
  const int head = blockIdx.y;
  const int wi = win / nww, wj = win % nww;
  const int i = threadIdx.x;  // token index inside the window: (i / ws, i % ws)
  const int hs = wi * ws + i / ws, wsft = wj * ws + i % ws;  // coordinates in the rolled image
  const int ph = (hs + shift) % H, pw = (wsft + shift) % W;  // source / destination pixel
  const int64_t pix = ((int64_t)b * H + ph) * W + pw;
  const float scale = rsqrtf((float)hd);
  {
    const float *src = qkv + pix * qkv_ps + head * hd;
    for (int dch = 0; dch < hd; ++dch) {
      sq[i * ldk + dch] = __fmul_rn(src[dch], scale);
      sk[i * ldk + dch] = src[C + dch];
      sv[i * hd + dch] = src[2 * C + dch];
    }
    int region = 0;
    if (shift > 0) {
      const int rh = hs < H - ws ? 0 : (hs < H - shift ? 1 : 2);
      const int rw = wsft < W - ws ? 0 : (wsft < W - shift ? 1 : 2);
      region = rh * 3 + rw;
    }
    s_region[i] = region;
  }
  __syncthreads();

  float s[T];
  const float *bias = rel_bias + ((int64_t)head * T + i) * T;
  const int my_region = s_region[i];
#pragma unroll
  for (int j = 0; j < T; ++j) s[j] = 0.f;
  for (int dch = 0; dch < hd; ++dch) {
    const float qv = sq[i * ldk + dch];
#pragma unroll
    for (int j = 0; j < T; ++j) s[j] = fmaf(qv, sk[j * ldk + dch], s[j]);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float v = s[j] + __ldg(bias + j);
    if (shift > 0 && s_region[j] != my_region) v += -100.0f;
    s[j] = v;
    mx = fmaxf(mx, v);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    s[j] = expf(s[j] - mx);
    sum += s[j];
  }
  const float inv = 1.0f / sum;
  float *dst = out + pix * out_ps + head * hd;
  for (int dch = 0; dch < hd; ++dch) {
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < T; ++j) o = fmaf(s[j] * inv, sv[j * hd + dch], o);
    dst[dch] = o;
  }
}

// Register-tiled variant for the head sizes the model uses (HD = 24 / 40 / 80): the thread's own q row lives in
// registers (loaded straight from global memory, 16 bytes at a time), K and V rows are read from shared memory as
// broadcast float4 (one LDS.128 feeds four FMAs — the scalar version above issues one LDS per FMA and is bound by the
// shared-memory pipe), and the output row is written with 16-byte stores.  The accumulation order of every dot
// product is unchanged, so results are bit-identical to the scalar kernel.
template <int T, int HD>
__global__ void __launch_bounds__(T)
window_attention_v4_kernel(const float *__restrict__ qkv, int qkv_ps, float *__restrict__ out, int out_ps,
                           const float *__restrict__ rel_bias, int H, int W, int C, int heads, int ws, int shift) {
  __shared__ __align__(16) float sk[T * HD];
  __shared__ __align__(16) float sv[T * HD];
  // the head's relative-position bias, staged through shared memory: thread i needs ROW i (bias[i][0..T)), and reading
  // it straight from global memory costs 32 L1 wavefronts per load instruction (rows are T floats apart) — measured:
  // those loads, not the FMAs, bounded the kernel.  Coalesced load + padded rows (T + 1) make both sides conflict free.
  __shared__ float s_bias[T * (T + 1)];
  __shared__ int s_region[T];
  const int nww = W / ws, nwh = H / ws;
  const int win = blockIdx.x % (nwh * nww);
  const int b = blockIdx.x / (nwh * nww);
  const int head = blockIdx.y;
  const int wi = win / nww, wj = win % nww;
  const int i = threadIdx.x;
  const int hs = wi * ws + i / ws, wsft = wj * ws + i % ws;
  const int ph = (hs + shift) % H, pw = (wsft + shift) % W;
  const int64_t pix = ((int64_t)b * H + ph) * W + pw;
  const float scale = rsqrtf((float)HD);
  float q[HD];
  {
    const float *src = qkv + pix * qkv_ps + head * HD;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 a = *reinterpret_cast<const float4 *>(src + 4 * c);
      const float4 k4 = *reinterpret_cast<const float4 *>(src + C + 4 * c);
      const float4 v4 = *reinterpret_cast<const float4 *>(src + 2 * C + 4 * c);
      q[4 * c] = __fmul_rn(a.x, scale); q[4 * c + 1] = __fmul_rn(a.y, scale);
      q[4 * c + 2] = __fmul_rn(a.z, scale); q[4 * c + 3] = __fmul_rn(a.w, scale);
      *reinterpret_cast<float4 *>(sk + i * HD + 4 * c) = k4;
      *reinterpret_cast<float4 *>(sv + i * HD + 4 * c) = v4;
    }
    int region = 0;
    if (shift > 0) {
      const int rh = hs < H - ws ? 0 : (hs < H - shift ? 1 : 2);
      const int rw = wsft < W - ws ? 0 : (wsft < W - shift ? 1 : 2);
      region = rh * 3 + rw;
    }
    s_region[i] = region;
    const float *hb = rel_bias + (int64_t)head * T * T;
#pragma unroll 8
    for (int e = i; e < T * T; e += T) s_bias[(e / T) * (T + 1) + (e % T)] = __ldg(hb + e);
  }
  __syncthreads();
  float s[T];
  const float *bias = s_bias + i * (T + 1);
  const int my_region = s_region[i];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 k4 = *reinterpret_cast<const float4 *>(sk + j * HD + 4 * c);  // broadcast read
      acc = fmaf(q[4 * c], k4.x, acc);
      acc = fmaf(q[4 * c + 1], k4.y, acc);
      acc = fmaf(q[4 * c + 2], k4.z, acc);
      acc = fmaf(q[4 * c + 3], k4.w, acc);
    }
    float v = acc + bias[j];
    if (shift > 0 && s_region[j] != my_region) v += -100.0f;
    s[j] = v;
    mx = fmaxf(mx, v);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    s[j] = expf(s[j] - mx);
    sum += s[j];
  }
  const float inv = 1.0f / sum;
  float o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const float p = s[j] * inv;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 v4 = *reinterpret_cast<const float4 *>(sv + j * HD + 4 * c);
      o[4 * c] = fmaf(p, v4.x, o[4 * c]);
      o[4 * c + 1] = fmaf(p, v4.y, o[4 * c + 1]);
      o[4 * c + 2] = fmaf(p, v4.z, o[4 * c + 2]);
      o[4 * c + 3] = fmaf(p, v4.w, o[4 * c + 3]);
    }
  }
  float *dst = out + pix * out_ps + head * HD;
#pragma unroll
  for (int c = 0; c < HD / 4; ++c)
    *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
}

template <int T, int HD>
int launch_attention_v4(const float *qkv, int qkv_ps, float *out, int out_ps, const float *rel_bias, int height, int width,
                        int channels, int heads, int window, int shift, dim3 grid, cudaStream_t st) {
  window_attention_v4_kernel<T, HD><<<grid, T, 0, st>>>(qkv, qkv_ps, out, out_ps, rel_bias, height, width, channels, heads,
                                                         window, shift);
  PCODEC_RETURN_LAUNCH();
}

// One thread per (output pixel, 4 consecutive patch entries): 32-bit index arithmetic (the 64-bit divisions of a
// thread-per-element version cost more than the memory traffic) and one 16-byte store per thread.
__global__ void __launch_bounds__(256)
im2col_nchw_kernel(const float *__restrict__ src, float *__restrict__ dst, int C, int H, int W, int k, int stride, int pad,
                   int OH, int OW, int k_pad, uint32_t total4) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  const uint32_t quads = (uint32_t)k_pad >> 2;
  const uint32_t pixel = t / quads, kq = t - pixel * quads;
  const uint32_t ow = pixel % (uint32_t)OW, r = pixel / (uint32_t)OW;
  const uint32_t oh = r % (uint32_t)OH, n = r / (uint32_t)OH;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = (int)kq * 4 + j;
    v[j] = 0.f;
    if (kk < k * k * C) {
      const int c = kk % C, tap = kk / C;
      const int ky = tap / k, kx = tap - ky * k;
      const int iy = (int)oh * stride + ky - pad, ix = (int)ow * stride + kx - pad;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[j] = __ldg(src + (((size_t)n * C + c) * H + iy) * (size_t)W + ix);
    }
  }
  *reinterpret_cast<float4 *>(dst + (size_t)pixel * k_pad + kq * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

}  // namespace

extern "C" int pcodec_window_attention(const float *qkv, int qkv_ps, float *out, int out_ps, const float *rel_bias,
                                       int batch, int height, int width, int channels, int heads, int window,
                                       int shift, void *stream) {
  if (!qkv || !out || !rel_bias || batch <= 0 || heads <= 0 || channels % heads != 0) return PCODEC_ERR_BAD_ARG;
  if (window <= 0 || height % window != 0 || width % window != 0 || shift < 0 || shift >= window)
    return PCODEC_ERR_BAD_ARG;
  const int T = window * window, hd = channels / heads;
  const size_t smem = sizeof(float) * ((size_t)2 * T * (hd + 1) + (size_t)T * hd);
  dim3 grid((unsigned)(batch * (height / window) * (width / window)), (unsigned)heads);
  cudaStream_t st = as_stream(stream);
  // 16-byte accesses need 16-byte aligned rows
  const bool vec_ok = (qkv_ps % 4 == 0) && (out_ps % 4 == 0) && (channels % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec_ok && T == 64 && hd == 24)
    return launch_attention_v4<64, 24>(qkv, qkv_ps, out, out_ps, rel_bias, height, width, channels, heads, window, shift, grid, st);
  if (vec_ok && T == 16 && hd == 40)
    return launch_attention_v4<16, 40>(qkv, qkv_ps, out, out_ps, rel_bias, height, width, channels, heads, window, shift, grid, st);
  if (vec_ok && T == 16 && hd == 80)
    return launch_attention_v4<16, 80>(qkv, qkv_ps, out, out_ps, rel_bias, height, width, channels, heads, window, shift, grid, st);
  if (T == 64) {
    if (smem > 48 * 1024)
      PCODEC_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
    window_attention_kernel<64><<<grid, 64, smem, st>>>(qkv, qkv_ps, out, out_ps, rel_bias, height, width, channels,
                                                        heads, window, shift);
  } else if (T == 16) {
    window_attention_kernel<16><<<grid, 16, smem, st>>>(qkv, qkv_ps, out, out_ps, rel_bias, height, width, channels,
                                                        heads, window, shift);
  } else {
    return PCODEC_ERR_UNSUPPORTED;
  }
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_im2col_nchw(const float *src, float *dst, int batch, int channels, int height, int width, int k,
                                  int stride, int pad, int out_h, int out_w, int k_pad, void *stream) {
  if (!src || !dst || batch <= 0 || channels <= 0 || k <= 0 || stride <= 0 || k_pad < k * k * channels)
    return PCODEC_ERR_BAD_ARG;
  if ((k_pad & 3) || (reinterpret_cast<uintptr_t>(dst) & 15)) return PCODEC_ERR_BAD_ARG;
  const int64_t total4 = (int64_t)batch * out_h * out_w * (k_pad / 4);
  if (total4 >= (1ll << 32)) return PCODEC_ERR_UNSUPPORTED;
  im2col_nchw_kernel<<<(unsigned)ceil_div64(total4, 256), 256, 0, as_stream(stream)>>>(
      src, dst, channels, height, width, k, stride, pad, out_h, out_w, k_pad, (uint32_t)total4);
  PCODEC_RETURN_LAUNCH();
}
