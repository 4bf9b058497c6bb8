// fp32 SIMT implicit-GEMM convolution ("sum of shifted-tap GEMMs") — the exact-fp32 path.
//
// One kernel covers every dense layer of the codec (reference: models/utils.py:186-204 conv/deconv,
// layers/layers.py:15-29 conv3x3/subpel/conv1x1, layers/gdn.py:50-63, win_attention.py qkv/proj Linear):
//   out[n, h*os+oy, w*os+ox, co] = epilogue( bias[co] + sum_t sum_ci in[n, h*is+dy_t, w*is+dx_t, ci] * W[t][ci][co] )
// A strided conv is (is = stride, os = 1); one phase of a stride-2 transposed conv is (is = 1, os = 2) with the
// taps of that phase; a Linear / 1x1 conv is a single (0,0) tap.  The input is a *virtual concatenation* of up
// to 4 NHWC segments, so the channel-conditional slice networks never materialise torch.cat support tensors.
//
// Tiling: CTA tile BM=128 rows x BN in {128,64,32,16} columns, BK=16; 256 threads as 16x16, each owning
// 8 x (BN/16) accumulators; double-buffered shared memory fed through registers so the global gather of
// K-step k+1 overlaps the FMAs of step k.  Accumulation order over K is fixed (segment, tap, channel), with
// no split-K and no atomics, so results are deterministic and independent of the batch size — the
// encoder and the decoder therefore reproduce each other's mu/sigma bit for bit.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int kThreads = 256;
constexpr int AS_STRIDE = BM + 4;

struct ConvParams {
  pcodec_conv_desc d;
  int64_t M;  // batch * grid_h * grid_w
};

template <int TN>
__device__ __forceinline__ void load_b_frag(const float *bs_row, int tn, float (&b)[TN]) {
  if constexpr (TN == 8) {
    const float4 v0 = *reinterpret_cast<const float4 *>(bs_row + tn * 4);
    const float4 v1 = *reinterpret_cast<const float4 *>(bs_row + 64 + tn * 4);
    b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
    b[4] = v1.x; b[5] = v1.y; b[6] = v1.z; b[7] = v1.w;
  } else if constexpr (TN == 4) {
    const float4 v0 = *reinterpret_cast<const float4 *>(bs_row + tn * 4);
    b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
  } else if constexpr (TN == 2) {
    const float2 v0 = *reinterpret_cast<const float2 *>(bs_row + tn * 2);
    b[0] = v0.x; b[1] = v0.y;
  } else {
    b[0] = bs_row[tn];
  }
}

// column of accumulator j for thread column tn
template <int TN>
__device__ __forceinline__ int col_of(int tn, int j) {
  if constexpr (TN == 8) return (j < 4) ? tn * 4 + j : 64 + tn * 4 + (j - 4);
  else return tn * TN + j;
}

__device__ __forceinline__ float apply_epilogue(int epi, float acc, float r1, float r2, bool has_r2) {
  switch (epi) {
    case PCODEC_EPI_GELU: return gelu_erf(acc);
    case PCODEC_EPI_ADD: return acc + r1;
    case PCODEC_EPI_ADD_GELU: return gelu_erf(acc + r1);
    case PCODEC_EPI_GATE: return r2 * sigmoid_f(acc) + r1;
    case PCODEC_EPI_GDN: return r1 * rsqrtf(acc);
    case PCODEC_EPI_IGDN: return r1 * sqrtf(acc);
    case PCODEC_EPI_LRP: {
      float v = __fadd_rn(r1, __fmul_rn(0.5f, tanhf(acc)));
      return has_r2 ? __fadd_rn(v, r2) : v;
    }
    case PCODEC_EPI_CLAMP01: return fminf(fmaxf(acc, 0.f), 1.f);
    case PCODEC_EPI_LEAKY: return acc > 0.f ? acc : __fmul_rn(0.01f, acc);
    case PCODEC_EPI_LEAKY_ADD: return (acc > 0.f ? acc : __fmul_rn(0.01f, acc)) + r1;
    default: return acc;
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads) conv_taps_simt_kernel(const __grid_constant__ ConvParams P) {
  constexpr int TN = BN / 16;
  constexpr int B_F4_PER_THREAD = (BK * BN / 4 + kThreads - 1) / kThreads;  // float4 loads of the weight tile
  __shared__ __align__(16) float As[2][BK][AS_STRIDE];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const pcodec_conv_desc &d = P.d;
  const int tid = threadIdx.x;
  const int tm = tid >> 4, tn = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A gather assignment: 2 rows per thread (row = tid/4 + 64*r), one 16-byte quad of the 16-channel chunk
  const int a_quad = tid & 3;
  int64_t a_img_base[2];  // n * in_h * in_w
  int a_h[2], a_w[2];
  bool a_row_ok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int64_t m = m0 + (tid >> 2) + 64 * r;
    a_row_ok[r] = m < P.M;
    const int64_t mm = a_row_ok[r] ? m : 0;
    const int gw = d.grid_w, gh = d.grid_h;
    const int w = (int)(mm % gw);
    const int64_t t = mm / gw;
    const int h = (int)(t % gh);
    const int64_t n = t / gh;
    a_img_base[r] = n * d.in_h * (int64_t)d.in_w;
    a_h[r] = h * d.in_step;
    a_w[r] = w * d.in_step;
  }
  const bool square = (d.flags & PCODEC_FLAG_SQUARE_INPUT) != 0;

  // ---- K iteration space: (segment, tap, 16-channel chunk)
  int seg = 0, tap = 0, chunk = 0;  // chunk counts 16-channel groups inside the segment
  int seg_cbase = 0;                // first concat channel of the current segment
  int64_t total_steps = 0;
  for (int s = 0; s < d.n_segments; ++s) total_steps += (int64_t)d.n_taps * (d.seg[s].channels / BK);

  float4 a_reg[2];
  float4 b_reg[B_F4_PER_THREAD];

  auto fetch = [&]() {
    // A
    const pcodec_segment &sg = d.seg[seg];
    const int dy = d.dy[tap], dx = d.dx[tap];
    const int c = chunk * BK + a_quad * 4;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int iy = a_h[r] + dy, ix = a_w[r] + dx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_row_ok[r] && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w) {
        const float *p = sg.ptr + (a_img_base[r] + (int64_t)iy * d.in_w + ix) * sg.pixel_stride + c;
        v = __ldg(reinterpret_cast<const float4 *>(p));
        if (square) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
      }
      a_reg[r] = v;
    }
    // B: rows k = 0..15 of W[tap][seg_cbase + chunk*16 + k][n0 .. n0+BN)
    const float *wbase = d.weight + ((int64_t)tap * d.cin_total + seg_cbase + chunk * BK) * d.cout + n0;
#pragma unroll
    for (int i = 0; i < B_F4_PER_THREAD; ++i) {
      const int f = tid + i * kThreads;  // float4 index inside the [BK][BN/4] tile
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < BK * BN / 4) {
        const int k = f / (BN / 4), nq = f % (BN / 4);
        const int col = n0 + nq * 4;
        const float *p = wbase + (int64_t)k * d.cout + nq * 4;
        if (col + 3 < d.cout && (d.cout & 3) == 0) {
          v = __ldg(reinterpret_cast<const float4 *>(p));
        } else {
          if (col + 0 < d.cout) v.x = __ldg(p + 0);
          if (col + 1 < d.cout) v.y = __ldg(p + 1);
          if (col + 2 < d.cout) v.z = __ldg(p + 2);
          if (col + 3 < d.cout) v.w = __ldg(p + 3);
        }
      }
      b_reg[i] = v;
    }
  };
  auto advance = [&]() {
    if (++chunk == d.seg[seg].channels / BK) {
      chunk = 0;
      if (++tap == d.n_taps) {
        tap = 0;
        seg_cbase += d.seg[seg].channels;
        ++seg;
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = (tid >> 2) + 64 * r;
      As[buf][a_quad * 4 + 0][row] = a_reg[r].x;
      As[buf][a_quad * 4 + 1][row] = a_reg[r].y;
      As[buf][a_quad * 4 + 2][row] = a_reg[r].z;
      As[buf][a_quad * 4 + 3][row] = a_reg[r].w;
    }
#pragma unroll
    for (int i = 0; i < B_F4_PER_THREAD; ++i) {
      const int f = tid + i * kThreads;
      if (f < BK * BN / 4) {
        const int k = f / (BN / 4), nq = f % (BN / 4);
        *reinterpret_cast<float4 *>(&Bs[buf][k][nq * 4]) = b_reg[i];
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  fetch();
  advance();
  stash(0);
  __syncthreads();

  for (int64_t step = 0; step < total_steps; ++step) {
    const int buf = (int)(step & 1);
    const bool more = step + 1 < total_steps;
    if (more) {
      fetch();
      advance();
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][tm * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + tm * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      load_b_frag<TN>(&Bs[buf][k][0], tn, b);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const bool shuffle = (d.flags & PCODEC_FLAG_PIXEL_SHUFFLE2) != 0;
  const bool has_r2 = d.r2 != nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = (i < 4) ? tm * 4 + i : 64 + tm * 4 + (i - 4);
    const int64_t m = m0 + row;
    if (m >= P.M) continue;
    const int w = (int)(m % d.grid_w);
    const int64_t t = m / d.grid_w;
    const int h = (int)(t % d.grid_h);
    const int64_t n = t / d.grid_h;
    const int oh = h * d.out_step + d.out_off_y, ow = w * d.out_step + d.out_off_x;
    const int64_t opix = (n * d.out_h + oh) * (int64_t)d.out_w + ow;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + col_of<TN>(tn, j);
      if (co >= d.cout) continue;
      float v = acc[i][j] + (d.bias ? __ldg(d.bias + co) : 0.f);
      if (!shuffle) {
        const float r1 = d.r1 ? d.r1[opix * d.r1_pixel_stride + co] : 0.f;
        const float r2 = d.r2 ? d.r2[opix * d.r2_pixel_stride + co] : 0.f;
        d.out[opix * d.out_pixel_stride + co] = apply_epilogue(d.epilogue, v, r1, r2, has_r2);
      } else {
        v = apply_epilogue(d.epilogue, v, 0.f, 0.f, false);
        const int c = co >> 2, si = (co >> 1) & 1, sj = co & 1;
        const int64_t sp = (n * d.out_h + (2 * oh + si)) * (int64_t)d.out_w + (2 * ow + sj);
        d.out[sp * d.out_pixel_stride + c] = v;
      }
    }
  }
}

template <int BN>
int launch_simt(const ConvParams &P, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(P.M, BM), (unsigned)((P.d.cout + BN - 1) / BN));
  conv_taps_simt_kernel<BN><<<grid, kThreads, 0, st>>>(P);
  PCODEC_RETURN_LAUNCH();
}

}  // namespace

int pcodec_conv_taps_tc(const pcodec_conv_desc *desc, void *stream);  // conv_tc.cu (tcgen05 path)
bool pcodec_conv_taps_tc_supported(const pcodec_conv_desc *desc);
int pcodec_conv_taps_tc16(const pcodec_conv_desc *desc, void *stream);  // conv_tc16.cu (fp16-split tcgen05 path)
bool pcodec_conv_taps_tc16_ready(const pcodec_conv_desc *desc);

static int validate(const pcodec_conv_desc *d) {
  if (!d || !d->weight) return PCODEC_ERR_BAD_ARG;
  const bool planes_out = (d->flags & PCODEC_FLAG_NO_F32_OUT) && d->out_hi && d->out_lo;  // fp16 path: no fp32 output
  if (!d->out && !planes_out) return PCODEC_ERR_BAD_ARG;
  if (d->n_segments < 1 || d->n_segments > PCODEC_MAX_SEGMENTS) return PCODEC_ERR_BAD_ARG;
  if (d->n_taps < 1 || d->n_taps > PCODEC_MAX_TAPS) return PCODEC_ERR_BAD_ARG;
  int cin = 0;
  for (int s = 0; s < d->n_segments; ++s) {
    const pcodec_segment &sg = d->seg[s];
    const bool planes_only = !sg.ptr && d->seg16[s].hi && d->seg16[s].lo;  // fp16 path: the segment exists as planes only
    if ((!sg.ptr && !planes_only) || sg.channels <= 0 || sg.channels % BK != 0 || sg.pixel_stride < sg.channels)
      return PCODEC_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(sg.ptr) & 15) || (sg.pixel_stride & 3)) return PCODEC_ERR_BAD_ARG;
    cin += sg.channels;
  }
  if (cin != d->cin_total || d->cout <= 0) return PCODEC_ERR_BAD_ARG;
  if (d->batch <= 0 || d->in_h <= 0 || d->in_w <= 0 || d->grid_h <= 0 || d->grid_w <= 0) return PCODEC_ERR_BAD_ARG;
  if (d->in_step < 1 || d->out_step < 1) return PCODEC_ERR_BAD_ARG;
  if ((d->flags & PCODEC_FLAG_PIXEL_SHUFFLE2) && (d->cout % 4 != 0)) return PCODEC_ERR_BAD_ARG;
  const int epi = d->epilogue;
  const bool needs_r1 = epi == PCODEC_EPI_ADD || epi == PCODEC_EPI_ADD_GELU || epi == PCODEC_EPI_GATE ||
                        epi == PCODEC_EPI_GDN || epi == PCODEC_EPI_IGDN || epi == PCODEC_EPI_LRP || epi == PCODEC_EPI_LEAKY_ADD;
  if (needs_r1 && !d->r1 && !d->r1_16.hi) return PCODEC_ERR_BAD_ARG;
  if (epi == PCODEC_EPI_GATE && !d->r2 && !d->r2_16.hi) return PCODEC_ERR_BAD_ARG;
  if ((d->flags & PCODEC_FLAG_PIXEL_SHUFFLE2) && needs_r1) return PCODEC_ERR_UNSUPPORTED;
  if ((d->flags & PCODEC_FLAG_SUBPIXEL_NCHW) && (needs_r1 || (d->flags & PCODEC_FLAG_PIXEL_SHUFFLE2) || d->out_step != 1))
    return PCODEC_ERR_UNSUPPORTED;
  return PCODEC_OK;
}

extern "C" int pcodec_conv_taps(const pcodec_conv_desc *desc, int impl, void *stream) {
  int rc = validate(desc);
  if (rc != PCODEC_OK) return rc;
  if (impl == 3 || (impl == 0 && pcodec_conv_taps_tc16_ready(desc))) return pcodec_conv_taps_tc16(desc, stream);
  // the fp32-input kernels below need every segment and the output as fp32
  if (!desc->out || (desc->flags & PCODEC_FLAG_NO_F32_OUT)) return PCODEC_ERR_UNSUPPORTED;
  for (int s = 0; s < desc->n_segments; ++s)
    if (!desc->seg[s].ptr) return PCODEC_ERR_UNSUPPORTED;
  if ((!desc->r1 && desc->r1_16.hi) || (!desc->r2 && desc->r2_16.hi)) return PCODEC_ERR_UNSUPPORTED;
  if (impl == 2 || (impl == 0 && pcodec_conv_taps_tc_supported(desc))) {
    if (!pcodec_conv_taps_tc_supported(desc)) return PCODEC_ERR_UNSUPPORTED;
    return pcodec_conv_taps_tc(desc, stream);
  }
  if (desc->flags & PCODEC_FLAG_SUBPIXEL_NCHW) return PCODEC_ERR_UNSUPPORTED;  // tcgen05 kernel only
  ConvParams P;
  P.d = *desc;
  P.M = (int64_t)desc->batch * desc->grid_h * desc->grid_w;
  cudaStream_t st = as_stream(stream);
  const int co = desc->cout;
  if (co <= 16) return launch_simt<16>(P, st);
  if (co <= 32) return launch_simt<32>(P, st);
  // prefer the widest tile that wastes < 1/8 of its columns
  auto waste = [&](int bn) { return (double)(((co + bn - 1) / bn) * bn - co) / co; };
  if (waste(128) <= 0.125) return launch_simt<128>(P, st);
  if (waste(64) <= 0.125) return launch_simt<64>(P, st);
  return launch_simt<32>(P, st);
}
