// fp16-split tcgen05 implicit-GEMM convolution: BOTH operands arrive by TMA, no register staging.
//
// Same contract as conv_tc.cu / conv_simt.cu (sum of shifted-tap GEMMs over a virtual channel concat), but
//   * activations are read from SPLIT-FP16 PLANES  x ~= hi + lo * 2^-11  (two fp16 NHWC tensors, 22 significant bits;
//     written by the producing kernel's epilogue or by split_planes_kernel), weights are pre-split the same way after a
//     per-layer power-of-two scale that keeps them out of the fp16 subnormals;
//   * the A operand (128 output pixels = a th x tw block of ONE image, 64 channels of one tap) is a 4-D TMA box of the
//     activation plane at the tap's shifted / strided coordinates: out-of-range pixels and channels are zero-filled by
//     the TMA unit, which IS the padding, the stride (elementStrides) and the ragged channel tail — no address math, no
//     loader or converter warps, no A ring in tensor memory;
//   * tcgen05.mma.kind::f16 (SS form, fp32 accumulate), three products per K step:
//         acc_hi[s % n_hi] += A_hi * B_hi        acc_lo += A_lo * B_hi + A_hi * B_lo     (acc_lo carries the 2^11)
//     i.e. half the tensor-pipe time per MAC of the 3xTF32 kernel and half the weight bytes per K step.  The tensor
//     core truncates once per MMA per accumulator, so the hi products round-robin over up to 4 accumulators chosen so
//     that none sees more than ~160 MMAs (conv_tc.cu measured that bound for fp32-class results);
//   * epilogue: (sum acc_hi + 2^-11 acc_lo) * 2^-wshift + bias, the fused epilogues of the other kernels, stores as fp32
//     NHWC and / or as split planes for the next convolution.
// The K order (segment, tap, 64-channel slab; hi*hi, lo*hi, hi*lo) is fixed: deterministic and batch invariant.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

using namespace pcodec_tc;

namespace {

constexpr int BM = 128;                // output pixels per tile (TMEM lanes)
constexpr int KS = 64;                 // fp16 channels per K slab = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * 128;      // one plane of one A tile
// warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2..: epilogue.  Two instantiations: EW = 8 epilogue warps
// (10 warps, <= 102 registers: two CTAs per SM for the short tiles) and EW = 12 (14 warps, one CTA per SM: a third warp
// per TMEM lane quarter shortens the un-overlapped epilogue of the 512-column tiles).
constexpr int EPI_WARPS_MAX = 12;
constexpr int SMEM_LIMIT = 225 * 1024;
constexpr int SMEM_HALF = 113 * 1024;  // two resident CTAs (tiles of <= 256 TMEM columns): 2 x (113 KB + 1 KB reserved) <= 228 KB
constexpr uint32_t PLAN_MAGIC = 0x31366370u;
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;

struct W16 {
  CUtensorMap map_hi, map_lo;
  __half *dev_hi, *dev_lo;  // [cout][k_stride]
  float *dev_scale;         // [2]: absmax bits scratch, 2^-wshift
  int cout, k_total, k_stride;
};

struct Plan16 {  // host side, stored in desc->plan
  uint32_t magic;
  int tw, th, tw_shift, tiles_w, tiles_h;
  int bn, n_tiles, stages, n_hi, n_lo, n_steps, n_ksteps, smem, ts;
  unsigned char maps[10 * sizeof(CUtensorMap)];  // a_hi[4] | a_lo[4] | w_hi | w_lo
};
static_assert(sizeof(Plan16) <= PCODEC_CONV_PLAN_BYTES, "plan scratch too small");

struct alignas(64) Params16 {
  CUtensorMap a_hi[PCODEC_MAX_SEGMENTS], a_lo[PCODEC_MAX_SEGMENTS], w_hi, w_lo;
  pcodec_conv_desc d;
  const float *w_scale;  // device: [1] = 2^-wshift
  int tw, th, tw_shift, tiles_w, tiles_h;
  int bn, stages, n_hi, n_lo, n_steps, n_ksteps;
  int ts;     // 1: the A slices go shared -> tensor memory with tcgen05.cp and the MMAs take A from TMEM (TS form)
  int debug;  // PCODEC_EXPERIMENTS builds only (timing experiments, wrong results): 1 = no A loads, 2 = no B loads, 4 = hi*hi MMAs only
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // low half = a
  return r;
}
__device__ __forceinline__ float h_lo_f(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))); }
__device__ __forceinline__ float h_hi_f(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }

// two fp32 -> packed (hi, hi) and (lo, lo) fp16 pairs of the split format
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  hi = pack_h2(a, b);
  float l0, l1;  // (a - hi(a)) * 2^11, both operations exact, as one packed subtract and one packed multiply
  f2_unpack(f2_mul(f2_sub(f2_pack(a, b), f2_pack(h_lo_f(hi), h_hi_f(hi))), f2_dup(LO_SCALE)), l0, l1);
  lo = pack_h2(l0, l1);
}

// ---------------------------------------------------------------------------------------------------------
// fp32 NHWC window -> split planes (8 channels per thread: two 16-byte loads, two 16-byte stores)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_planes_kernel(const float *__restrict__ src, int src_ps, int64_t n_pixels, int channels, uint16_t *__restrict__ hi,
                    uint16_t *__restrict__ lo, int dst_ps, int square) {
  const int groups = channels >> 3;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pixels * groups) return;
  const int64_t pix = t / groups;
  const int g = (int)(t - pix * groups);
  const float4 *s = reinterpret_cast<const float4 *>(src + pix * src_ps + 8 * g);
  float4 a = __ldg(s), b = __ldg(s + 1);
  if (square) {
    const float k = 0.0625f;  // (x * 2^-4)^2: keeps |x| up to 4095 inside the fp16 range; the GDN launch restores the 2^8
    a.x *= k; a.y *= k; a.z *= k; a.w *= k; b.x *= k; b.y *= k; b.z *= k; b.w *= k;
    a.x *= a.x; a.y *= a.y; a.z *= a.z; a.w *= a.w; b.x *= b.x; b.y *= b.y; b.z *= b.z; b.w *= b.w;
  }
  uint4 h, l;
  split2(a.x, a.y, h.x, l.x);
  split2(a.z, a.w, h.y, l.y);
  split2(b.x, b.y, h.z, l.z);
  split2(b.z, b.w, h.w, l.w);
  *reinterpret_cast<uint4 *>(hi + pix * dst_ps + 8 * g) = h;
  *reinterpret_cast<uint4 *>(lo + pix * dst_ps + 8 * g) = l;
}

// ---------------------------------------------------------------------------------------------------------
// weights: per-layer absmax -> power-of-two scale -> fp16 hi / lo [cout][k_stride]
// ---------------------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float *__restrict__ w, int64_t n, float *__restrict__ scratch) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int *>(scratch), __float_as_uint(m));  // m >= 0
}

__global__ void split_weights16_kernel(const float *__restrict__ w_tap_major, int n_taps, int cin, int cout, int k_stride,
                                       float *__restrict__ scale, __half *__restrict__ hi, __half *__restrict__ lo) {
  // in: [tap][cin][cout]; out: [cout][k_stride] (k = tap*cin + ci, zero padded to k_stride)
  const float amax = scale[0];
  // 2^wshift * amax in [2^13, 2^14): the largest weight sits well inside the fp16 range, and a weight 2^-13 of it
  // still has a NORMAL lo part
  int e = 0;
  if (amax > 0.f) frexpf(amax, &e);  // amax = m * 2^e, m in [0.5, 1)
  const int wshift = 14 - e;
  const float up = ldexpf(1.0f, wshift);
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[1] = ldexpf(1.0f, -wshift);
  const int64_t total = (int64_t)cout * k_stride;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % k_stride), co = (int)(i / k_stride);
  float w = 0.f;
  if (k < n_taps * cin) w = w_tap_major[(int64_t)k * cout + co] * up;
  const __half h = __float2half_rn(w);
  hi[i] = h;
  lo[i] = __float2half_rn((w - __half2float(h)) * LO_SCALE);
}

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
#ifdef PCODEC_EXPERIMENTS
// timeline of one CTA (PCODEC_TC16_DEBUG bit 6; bit 7: a mid-grid CTA): clock64 at
//   0 kernel entry, 1 setup done (barriers, TMEM), 2 first stage landed, 3 last MMA committed (issuer), 4 accumulators
//   complete (epilogue warps released), 5 epilogue done, 6 kernel exit; 8+s: MMA issue of slab s (s < 100)
__device__ long long g_trace16[128];
#define T16(ev) do { if (trace) g_trace16[ev] = clock64(); } while (0)
#else
#define T16(ev) do { } while (0)
#endif

template <int EW>
__global__ void __launch_bounds__(32 * (EW + 2), EW == 8 ? 2 : 1)
conv_taps_tc16_kernel(const __grid_constant__ Params16 P) {
  constexpr int EPI_WARPS = EW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const pcodec_conv_desc &d = P.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, stages = P.stages, n_steps = P.n_steps;
  const int b_bytes = bn * 128;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;  // A_hi | A_lo | B_hi | B_lo

  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (what the 128-byte swizzle
  // atoms need); an alignment pad would cost the 2-stage configuration of the two-CTA tiles its second CTA
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  auto a_hi = [&](int s) { return smem_base + s * stage_bytes; };
  auto a_lo = [&](int s) { return smem_base + s * stage_bytes + A_BYTES; };
  auto b_hi = [&](int s) { return smem_base + s * stage_bytes + 2 * A_BYTES; };
  auto b_lo = [&](int s) { return smem_base + s * stage_bytes + 2 * A_BYTES + b_bytes; };
  // the stage area doubles as the epilogue's transpose scratch (2 KB per epilogue warp)
  const uint32_t bar_base = smem_base + max(stages * stage_bytes, EPI_WARPS * 2048);
  auto full = [&](int s) { return bar_base + 8u * s; };              // TMA bytes of the stage landed
  auto empty = [&](int s) { return bar_base + 8u * (stages + s); };  // MMAs of the stage retired (tcgen05.commit)
  const uint32_t tmem_full = bar_base + 8u * (2 * stages);
  const uint32_t tmem_slot = tmem_full + 8u;
  uint8_t *smem_gen = smem_raw;

#ifdef PCODEC_EXPERIMENTS
  const bool trace = (P.debug & 64) && blockIdx.x == ((P.debug & 128) ? (gridDim.x / 2) & ~7u : 0) && lane == 0;
#endif
  if (warp == 2) T16(0);
  const int n_acc = P.n_hi + P.n_lo;  // hi accumulators + lo accumulators, both used round-robin over the K steps
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_acc * bn + (P.ts ? 64 : 0)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    // the TMA unit fetches each 128-byte tensor map on first use: start those fetches before anything else
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.a_hi[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.a_lo[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.w_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.w_lo) : "memory");
    for (int s = 0; s < stages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  // tcgen05.alloc is .sync.aligned: the WHOLE warp must arrive converged, so it runs in a warp that has not diverged
  // (the barrier initialisation by a single lane lives in warp 0 for that reason)
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
  if (warp == 2) T16(1);

  // tile = th x tw block of output-grid pixels of one image
  // 1-D grid, N tile fastest: the CTAs that share an A tile (same pixels, different output channels) are scheduled
  // together, so the activation planes are read from DRAM once instead of once per N tile (ncu: 1.81x -> of the
  // algorithmic bytes on the 5x5 stride-2 layer when the N tile was the slow index)
  const int n_tiles = (d.cout + bn - 1) / bn;
  int t = blockIdx.x / n_tiles;
  const int n_tile = blockIdx.x - t * n_tiles;
  const int twi = t % P.tiles_w;
  t /= P.tiles_w;
  const int thi = t % P.tiles_h;
  const int img = t / P.tiles_h;
  const int h0 = thi * P.th, w0 = twi * P.tw;
  const int n0 = n_tile * bn;

  if (warp == 0) {
    // =============================== TMA producer (A planes + weights) ===============================
    if (elect_one()) {
      int seg = 0, tap = 0, kc = 0, seg_cbase = 0, st = 0;
      uint32_t ph = 1;
      for (int s = 0; s < n_steps; ++s) {
        mbar_wait(empty(st), ph);
        const int c = kc * KS;
        const int x = w0 * d.in_step + d.dx[tap], y = h0 * d.in_step + d.dy[tap];
        const int k = tap * d.cin_total + seg_cbase + c;
#ifdef PCODEC_EXPERIMENTS
        mbar_expect_tx(full(st), (uint32_t)(((P.debug & 1) ? 0 : 2 * A_BYTES) + ((P.debug & 2) ? 0 : 2 * b_bytes)));
        if (!(P.debug & 1)) {
          tma_load_4d(a_hi(st), &P.a_hi[seg], full(st), c, x, y, img);
          tma_load_4d(a_lo(st), &P.a_lo[seg], full(st), c, x, y, img);
        }
        if (!(P.debug & 2)) {
          tma_load_2d(b_hi(st), &P.w_hi, full(st), k, n0);
          tma_load_2d(b_lo(st), &P.w_lo, full(st), k, n0);
        }
#else
        mbar_expect_tx(full(st), (uint32_t)stage_bytes);
        tma_load_4d(a_hi(st), &P.a_hi[seg], full(st), c, x, y, img);
        tma_load_4d(a_lo(st), &P.a_lo[seg], full(st), c, x, y, img);
        tma_load_2d(b_hi(st), &P.w_hi, full(st), k, n0);
        tma_load_2d(b_lo(st), &P.w_lo, full(st), k, n0);
#endif
        const int sc = d.seg[seg].channels;
        if (++kc == (sc + KS - 1) / KS) {
          kc = 0;
          if (++tap == d.n_taps) { tap = 0; seg_cbase += sc; ++seg; }
        }
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // instruction descriptor: D = f32 (1 << 4), A = B = f16 (format 0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    // The K steps round-robin over n_hi hi and n_lo lo accumulators: every accumulator sees a fixed sequence of products
    // (deterministic), and the epilogue adds them up with round-to-nearest.  (Several lo accumulators were tried against
    // the idea that back-to-back MMAs into one accumulator serialise: no change, so n_lo = 1 by default.)
    const uint32_t lo_base = tmem_acc + (uint32_t)(P.n_hi * bn);
    const uint32_t a_tmem = tmem_acc + (uint32_t)(n_acc * bn);  // TS form: 4 K steps x (hi 8 + lo 8) columns
    if (elect_one()) {
      int seg = 0, tap = 0, kc = 0, st = 0, hi_idx = 0, lo_idx = 0;
      uint32_t ph = 0, hi_init = 0, lo_init = 0;
      const int n_hi = P.n_hi, n_lo = P.n_lo;
      for (int s = 0; s < n_steps; ++s) {
        const int sc = d.seg[seg].channels;
        const int ksteps = (min(KS, sc - kc * KS) + 15) >> 4;  // K steps of 16 channels that hold data (the rest is zero fill)
        mbar_wait(full(st), ph);
        tc_fence_after();
#ifdef PCODEC_EXPERIMENTS
        if (s == 0) T16(2);
        if (s < 100) T16(8 + s);
#endif
        const uint64_t da_hi = umma_desc_sw128(a_hi(st)), da_lo = umma_desc_sw128(a_lo(st));
        const uint64_t db_hi = umma_desc_sw128(b_hi(st)), db_lo = umma_desc_sw128(b_lo(st));
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);  // 16 fp16 = 32 bytes = 2 x 16-byte units inside the swizzle row
          if (P.ts) {
            // The SS form re-reads the 128-row A slice from shared memory for every MMA: 32 bytes out of each 128-byte row,
            // ~100 clk of shared-memory wavefronts whatever N — more than the N/2 clk of a narrow tile's MMA (measured:
            // 83 clk per MMA at N = 112, 107 at N = 32).  Copy each slice ONCE into tensor memory and use the TS form.
            const uint32_t ta_hi = a_tmem + (uint32_t)(k * 16), ta_lo = ta_hi + 8u;
            tmem_cp_128x256b(ta_hi, da_hi + adv);
            tmem_cp_128x256b(ta_lo, da_lo + adv);
            umma_f16_ts(tmem_acc + (uint32_t)(hi_idx * bn), ta_hi, db_hi + adv, idesc, (hi_init >> hi_idx) & 1u);
            hi_init |= 1u << hi_idx;
            if (++hi_idx == n_hi) hi_idx = 0;
            umma_f16_ts(lo_base + (uint32_t)(lo_idx * bn), ta_lo, db_hi + adv, idesc, (lo_init >> lo_idx) & 1u);
            lo_init |= 1u << lo_idx;
            if (++lo_idx == n_lo) lo_idx = 0;
            umma_f16_ts(lo_base + (uint32_t)(lo_idx * bn), ta_hi, db_lo + adv, idesc, (lo_init >> lo_idx) & 1u);
            lo_init |= 1u << lo_idx;
            if (++lo_idx == n_lo) lo_idx = 0;
            continue;
          }
          umma_f16_ss(tmem_acc + (uint32_t)(hi_idx * bn), da_hi + adv, db_hi + adv, idesc, (hi_init >> hi_idx) & 1u);
          hi_init |= 1u << hi_idx;
          if (++hi_idx == n_hi) hi_idx = 0;
#ifdef PCODEC_EXPERIMENTS
          if (P.debug & 4) continue;
#endif
          umma_f16_ss(lo_base + (uint32_t)(lo_idx * bn), da_lo + adv, db_hi + adv, idesc, (lo_init >> lo_idx) & 1u);
          lo_init |= 1u << lo_idx;
          if (++lo_idx == n_lo) lo_idx = 0;
          umma_f16_ss(lo_base + (uint32_t)(lo_idx * bn), da_hi + adv, db_lo + adv, idesc, (lo_init >> lo_idx) & 1u);
          lo_init |= 1u << lo_idx;
          if (++lo_idx == n_lo) lo_idx = 0;
        }
        umma_commit(empty(st));
        if (++kc == (sc + KS - 1) / KS) {
          kc = 0;
          if (++tap == d.n_taps) { tap = 0; ++seg; }
        }
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full);
      T16(3);
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ===============================
    const int ew = warp - 2;            // 0..7
    const int quarter = warp & 3;       // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int half = ew >> 2;           // the EW/4 warps of a quarter interleave the 16-column groups
    constexpr int CSTEP = 16 * (EW / 4);  // columns between two groups of the same warp
    const int row = quarter * 32 + lane;
    const int ti = row >> P.tw_shift, tj = row & (P.tw - 1);
    const int gh = h0 + ti, gw = w0 + tj;
    const bool row_ok = gh < d.grid_h && gw < d.grid_w;
    const int oh = gh * d.out_step + d.out_off_y, ow = gw * d.out_step + d.out_off_x;
    const int64_t opix = row_ok ? ((int64_t)img * d.out_h + oh) * (int64_t)d.out_w + ow : 0;
    const bool shuffle = (d.flags & PCODEC_FLAG_PIXEL_SHUFFLE2) != 0;
    const bool has_r2 = d.r2 != nullptr;
    const bool want_f32 = !(d.flags & PCODEC_FLAG_NO_F32_OUT);
    const bool want_planes = d.out_hi != nullptr;

    const bool r1p = !d.r1 && d.r1_16.hi, r2p = !d.r2 && d.r2_16.hi;  // residuals given as split planes
    if ((d.r1 || d.r2 || r1p || r2p) && !shuffle && row_ok) {
      // residual values do not depend on the accumulators: pull them into L2 while the main loop runs
      for (int c = half * 32; c < bn && n0 + c < d.cout; c += 32 * (EW / 4)) {
        if (d.r1) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r1 + opix * d.r1_pixel_stride + n0 + c));
        if (d.r2) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r2 + opix * d.r2_pixel_stride + n0 + c));
      }
      for (int c = half * 64; c < bn && n0 + c < d.cout; c += 64 * (EW / 4)) {
        if (r1p) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r1_16.hi + opix * d.r1_16.pixel_stride + n0 + c));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r1_16.lo + opix * d.r1_16.pixel_stride + n0 + c));
        }
        if (r2p) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r2_16.hi + opix * d.r2_16.pixel_stride + n0 + c));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r2_16.lo + opix * d.r2_16.pixel_stride + n0 + c));
        }
      }
    }
    // residual operand for 4 channels of one pixel: fp32 when given, else reconstructed from its planes
    auto load_res = [&](const float *r, int rps, const pcodec_planes &pl, bool from_planes, int64_t pix, int co, bool ok) {
      if (!ok) return make_float4(0.f, 0.f, 0.f, 0.f);
      if (r) return __ldg(reinterpret_cast<const float4 *>(r + pix * rps + co));
      if (!from_planes) return make_float4(0.f, 0.f, 0.f, 0.f);
      const uint2 h = __ldg(reinterpret_cast<const uint2 *>(pl.hi + pix * pl.pixel_stride + co));
      const uint2 l = __ldg(reinterpret_cast<const uint2 *>(pl.lo + pix * pl.pixel_stride + co));
      return make_float4(__fmaf_rn(h_lo_f(l.x), LO_INV, h_lo_f(h.x)), __fmaf_rn(h_hi_f(l.x), LO_INV, h_hi_f(h.x)),
                         __fmaf_rn(h_lo_f(l.y), LO_INV, h_lo_f(h.y)), __fmaf_rn(h_hi_f(l.y), LO_INV, h_hi_f(h.y)));
    };
    // first residual in two steps (16 raw bytes per row now, the values later) so that its loads can be issued early
    const bool has_r1 = d.r1 || r1p;
    auto res_raw = [&](int64_t pix, int co, bool ok) {
      if (!ok) return make_uint4(0u, 0u, 0u, 0u);
      if (d.r1) return __ldg(reinterpret_cast<const uint4 *>(d.r1 + pix * d.r1_pixel_stride + co));
      const uint2 h = __ldg(reinterpret_cast<const uint2 *>(d.r1_16.hi + pix * d.r1_16.pixel_stride + co));
      const uint2 l = __ldg(reinterpret_cast<const uint2 *>(d.r1_16.lo + pix * d.r1_16.pixel_stride + co));
      return make_uint4(h.x, h.y, l.x, l.y);
    };
    auto res_val = [&](const uint4 &w) {
      if (d.r1) return make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
      return make_float4(__fmaf_rn(h_lo_f(w.z), LO_INV, h_lo_f(w.x)), __fmaf_rn(h_hi_f(w.z), LO_INV, h_hi_f(w.x)),
                         __fmaf_rn(h_lo_f(w.w), LO_INV, h_lo_f(w.y)), __fmaf_rn(h_hi_f(w.w), LO_INV, h_hi_f(w.y)));
    };
    const bool square_planes = (d.flags & PCODEC_FLAG_SQUARE_OUT_PLANES) != 0;
    const float out_scale = __ldg(P.w_scale + 1) * ((d.flags & PCODEC_FLAG_SQUARE_INPUT) ? 256.0f : 1.0f);

    // Every lane polls the barrier in its own (inline-asm) loop, so lanes may leave it in different iterations and the
    // warp is NOT guaranteed to be converged afterwards — but everything below is .sync.aligned (tcgen05.ld), which
    // requires the whole warp to execute it together.  Re-converge explicitly.  (Without this the kernel ran correctly
    // almost always and faulted once in a few thousand launches, depending on timing: the intermittent device fault
    // of round 1.)
    mbar_wait(tmem_full, 0);
    __syncwarp();
    tc_fence_after();
    if (warp == 2) T16(4);
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const int n_hi_used = min(P.n_hi, P.n_ksteps);      // very short reductions touch fewer accumulators
    const int n_lo_used = min(P.n_lo, 2 * P.n_ksteps);
    // acc[16] <- (sum of the hi accumulators + 2^-11 * sum of the lo accumulators) * 2^-wshift for columns c0 .. c0+15
    auto load_acc = [&](int c0, float (&acc)[16]) {
      uint32_t t0[16], t1[16];
      tmem_ld16_issue(lane_addr + (uint32_t)c0, t0);
      tmem_ld16_issue(lane_addr + (uint32_t)(P.n_hi * bn + c0), t1);
      tmem_wait_ld();
      tmem_pin(t0);
      tmem_pin(t1);
      if (n_hi_used == 1 && n_lo_used == 1) {
        // the common case of the epilogue-bound layers (short reductions): no accumulator sums, and the combination
        // runs on packed pairs straight out of the registers the TMEM loads filled (same operations per element)
        const F2 inv = f2_dup(LO_INV), sc = f2_dup(out_scale);
#pragma unroll
        for (int j = 0; j < 16; j += 2)
          f2_unpack(f2_mul(f2_fma(f2_pack_bits(t1[j], t1[j + 1]), inv, f2_pack_bits(t0[j], t0[j + 1])), sc), acc[j], acc[j + 1]);
        return;
      }
      float lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc[j] = __uint_as_float(t0[j]); lo[j] = __uint_as_float(t1[j]); }
#pragma unroll 1
      for (int a = 1; a < max(n_hi_used, n_lo_used); ++a) {
        const bool h = a < n_hi_used, l = a < n_lo_used;  // warp-uniform
        if (h) tmem_ld16_issue(lane_addr + (uint32_t)(a * bn + c0), t0);
        if (l) tmem_ld16_issue(lane_addr + (uint32_t)((P.n_hi + a) * bn + c0), t1);
        tmem_wait_ld();
        if (h) {
          tmem_pin(t0);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(t0[j]);
        }
        if (l) {
          tmem_pin(t1);
#pragma unroll
          for (int j = 0; j < 16; ++j) lo[j] += __uint_as_float(t1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __fmul_rn(__fmaf_rn(lo[j], LO_INV, acc[j]), out_scale);
    };

    if (d.flags & PCODEC_FLAG_SUBPIXEL_NCHW) {
      // Image layer: conv channel (2*py + px) * C + c  ->  out[n][c][2h + py][2w + px] (NCHW).  cout == 16: one column
      // group, done by the first warp of each quarter.
      const int Cimg = d.out_pixel_stride;
      if (half == 0) {
        float acc[16];
        load_acc(0, acc);
        if (row_ok) {
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = d.bias ? __ldg(reinterpret_cast<const float4 *>(d.bias) + j4) : z4;
            const float4 o = tc_epilogue4(d.epilogue, make_float4(acc[4 * j4] + b4.x, acc[4 * j4 + 1] + b4.y,
                                                                   acc[4 * j4 + 2] + b4.z, acc[4 * j4 + 3] + b4.w), z4, z4, false);
            acc[4 * j4] = o.x; acc[4 * j4 + 1] = o.y; acc[4 * j4 + 2] = o.z; acc[4 * j4 + 3] = o.w;
          }
          const int64_t plane = (int64_t)d.out_h * d.out_w;
          for (int c = 0; c < Cimg && c < 4; ++c)
            for (int py = 0; py < 2; ++py) {
              float v0 = 0.f, v1 = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (j == (2 * py) * Cimg + c) v0 = acc[j];
                if (j == (2 * py + 1) * Cimg + c) v1 = acc[j];
              }
              float *dst = d.out + ((int64_t)img * Cimg + c) * plane + (int64_t)(2 * gh + py) * d.out_w + 2 * gw;
              *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
            }
        }
      }
    } else if (!shuffle) {
      // Coalesced epilogue: each warp transposes its 32 rows x 16 columns through a private 2 KB shared-memory tile
      // (the pipeline stages are idle by now), then works with 4 lanes per row / 8 rows per instruction: residual loads
      // and fp32 stores are 64-byte runs, plane stores 32-byte runs.
      uint8_t *stg = smem_gen + ew * 2048;
      const int rl = lane >> 2, cc = lane & 3;
      int64_t opix_t[4];
      uint32_t ok_t = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int src = i * 8 + rl;
        const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)opix, src);
        const uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)((uint64_t)opix >> 32), src);
        opix_t[i] = (int64_t)(((uint64_t)hi << 32) | lo);
        ok_t |= (__shfl_sync(0xFFFFFFFFu, row_ok ? 1u : 0u, src) & 1u) << i;
      }
      const uint32_t wr_off = (uint32_t)lane * 64u, wr_x = (uint32_t)(lane >> 1) & 3u;
      for (int c0 = half * 16; c0 < bn; c0 += CSTEP) {
        if (n0 + c0 >= d.cout) break;  // padded last N tile
        const int co = n0 + c0 + 4 * cc;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 bias4 = d.bias ? __ldg(reinterpret_cast<const float4 *>(d.bias + co)) : z4;
        const F2 bias_xy = f2_pack(bias4.x, bias4.y), bias_zw = f2_pack(bias4.z, bias4.w);
        float acc[16];
        load_acc(c0, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4 *>(stg + wr_off + (((uint32_t)q ^ wr_x) << 4)) =
              make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        // Residual rows are fetched ahead of their use (ncu: a fifth of the epilogue warps' stall samples sat on the first
        // use of these loads): pair (0, 1) before the transposition is complete, pair (2, 3) before pair (0, 1) is
        // processed.  16 raw bytes per row; the values are formed at the point of use.  (Fetching before the TMEM loads
        // instead cost registers across them and was slower; an L1 prefetch of the rows was slower too.)
        uint4 wq[2][2];
        wq[0][0] = wq[0][1] = wq[1][0] = wq[1][1] = make_uint4(0u, 0u, 0u, 0u);
        if (has_r1) {
          wq[0][0] = res_raw(opix_t[0], co, ok_t & 1u);
          wq[0][1] = res_raw(opix_t[1], co, (ok_t >> 1) & 1u);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; i += 2) {  // two tile rows (i, i + 1) per call
          F8 v, a1, a2;
          const int r0 = i * 8 + rl, r1_ = (i + 1) * 8 + rl;
          v.a = *reinterpret_cast<const float4 *>(stg + r0 * 64 + (((uint32_t)cc ^ ((uint32_t)(r0 >> 1) & 3u)) << 4));
          v.b = *reinterpret_cast<const float4 *>(stg + r1_ * 64 + (((uint32_t)cc ^ ((uint32_t)(r1_ >> 1) & 3u)) << 4));
          const bool ok0 = (ok_t >> i) & 1u, ok1 = (ok_t >> (i + 1)) & 1u;
          if (i == 0 && has_r1) {  // the second row pair's loads fly while the first pair is processed
            wq[1][0] = res_raw(opix_t[2], co, (ok_t >> 2) & 1u);
            wq[1][1] = res_raw(opix_t[3], co, (ok_t >> 3) & 1u);
          }
          a1.a = a1.b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_r1) { a1.a = res_val(wq[i >> 1][0]); a1.b = res_val(wq[i >> 1][1]); }
          a2.a = load_res(d.r2, d.r2_pixel_stride, d.r2_16, r2p, opix_t[i], co, ok0);
          a2.b = load_res(d.r2, d.r2_pixel_stride, d.r2_16, r2p, opix_t[i + 1], co, ok1);
          f2_unpack(f2_add(f2_pack(v.a.x, v.a.y), bias_xy), v.a.x, v.a.y);
          f2_unpack(f2_add(f2_pack(v.a.z, v.a.w), bias_zw), v.a.z, v.a.w);
          f2_unpack(f2_add(f2_pack(v.b.x, v.b.y), bias_xy), v.b.x, v.b.y);
          f2_unpack(f2_add(f2_pack(v.b.z, v.b.w), bias_zw), v.b.z, v.b.w);
          const F8 o = tc_epilogue8(d.epilogue, v, a1, a2, has_r2);
          if (want_f32) {
            if (ok0) *reinterpret_cast<float4 *>(d.out + opix_t[i] * d.out_pixel_stride + co) = o.a;
            if (ok1) *reinterpret_cast<float4 *>(d.out + opix_t[i + 1] * d.out_pixel_stride + co) = o.b;
          }
          if (want_planes) {
            uint2 h, l;
            F8 p = o;
            if (square_planes) {  // the GDN that follows reads (x * 2^-4)^2
              p.a.x *= 0.0625f; p.a.y *= 0.0625f; p.a.z *= 0.0625f; p.a.w *= 0.0625f;
              p.b.x *= 0.0625f; p.b.y *= 0.0625f; p.b.z *= 0.0625f; p.b.w *= 0.0625f;
              p.a.x *= p.a.x; p.a.y *= p.a.y; p.a.z *= p.a.z; p.a.w *= p.a.w;
              p.b.x *= p.b.x; p.b.y *= p.b.y; p.b.z *= p.b.z; p.b.w *= p.b.w;
            }
            if (ok0) {
              split2(p.a.x, p.a.y, h.x, l.x);
              split2(p.a.z, p.a.w, h.y, l.y);
              *reinterpret_cast<uint2 *>(d.out_hi + opix_t[i] * d.out_plane_stride + co) = h;
              *reinterpret_cast<uint2 *>(d.out_lo + opix_t[i] * d.out_plane_stride + co) = l;
            }
            if (ok1) {
              split2(p.b.x, p.b.y, h.x, l.x);
              split2(p.b.z, p.b.w, h.y, l.y);
              *reinterpret_cast<uint2 *>(d.out_hi + opix_t[i + 1] * d.out_plane_stride + co) = h;
              *reinterpret_cast<uint2 *>(d.out_lo + opix_t[i + 1] * d.out_plane_stride + co) = l;
            }
          }
        }
        __syncwarp();
      }
    } else {
      // pixel shuffle: 4 consecutive conv channels = the 2x2 sub-pixels of one output channel (subpel_conv3x3)
      for (int c0 = half * 16; c0 < bn; c0 += CSTEP) {
        if (n0 + c0 >= d.cout) break;
        __syncwarp();  // lanes that skipped the stores of the previous round (`continue`) rejoin before the collective load
        float acc[16];
        load_acc(c0, acc);  // warp-collective
        if (!row_ok) continue;
        const int co0 = n0 + c0;
#pragma unroll 1
        for (int j4 = 0; j4 < 4; ++j4) {
          const int co = co0 + 4 * j4;
          const float4 b4 = d.bias ? __ldg(reinterpret_cast<const float4 *>(d.bias + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float a4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) a4[e] = acc[0];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            if ((jj >> 2) == j4) a4[jj & 3] = acc[jj];
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 o = tc_epilogue4(d.epilogue, make_float4(a4[0] + b4.x, a4[1] + b4.y, a4[2] + b4.z, a4[3] + b4.w), z, z, false);
          const int c = co >> 2;
          const int64_t sp0 = ((int64_t)img * d.out_h + 2 * oh) * (int64_t)d.out_w + 2 * ow;
          const int64_t sp[4] = {sp0, sp0 + 1, sp0 + d.out_w, sp0 + d.out_w + 1};
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (want_f32) d.out[sp[e] * d.out_pixel_stride + c] = ov[e];
            if (want_planes) {
              const __half hh = __float2half_rn(ov[e]);
              d.out_hi[sp[e] * d.out_plane_stride + c] = __half_as_ushort(hh);
              d.out_lo[sp[e] * d.out_plane_stride + c] = __half_as_ushort(__float2half_rn((ov[e] - __half2float(hh)) * LO_SCALE));
            }
          }
        }
      }
    }
    tc_fence_before();
    if (warp == 2) T16(5);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
  if (warp == 2) T16(6);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

int max_mma_per_acc() {
  static const int v = [] {
    const char *e = pcodec_knob("PCODEC_TC16_MAX_MMA");
    const int n = e ? atoi(e) : 0;
    return n > 0 ? n : 160;
  }();
  return v;
}

// N tile: the fewest tiles (least A re-reading) such that enough hi accumulators fit the 512 TMEM columns next to the lo
// accumulator for no hi accumulator to see more than max_mma_per_acc() MMAs.  Short reductions (<= 24 hi MMAs, i.e. the
// 1x1 layers) spend most of a tile's life in its serial prologue / epilogue: they get tiles of <= 128 columns with ONE hi
// accumulator (<= 256 TMEM columns) so that two CTAs share an SM and overlap each other.
int pick_bn16(int cout, int n_mma_hi, int *n_hi_out, int *n_lo_out, int ts) {
  if (cout % 16 != 0) return 0;
  const int cols_full = ts ? 448 : 512, cols_pair = ts ? 192 : 256;  // TS form: 64 columns hold the A slices
  static const bool pair_short = [] { const char *e = pcodec_knob("PCODEC_TC16_PAIR"); return !e || atoi(e) != 0; }();
  // one lo accumulator by default: a second one measured no faster (back-to-back MMAs into one accumulator are not the
  // bound) and costs the long reductions a hi accumulator, i.e. accuracy; PCODEC_TC16_NLO keeps the experiment reachable
  static const int lo_max = [] { const char *e = pcodec_knob("PCODEC_TC16_NLO"); return e ? std::max(1, std::min(4, atoi(e))) : 1; }();
  auto finish = [&](int bn, int cols, int need) {
    // spend the TMEM columns next to `need` hi accumulators on lo accumulators (2 break the lo -> lo dependency, narrow
    // tiles want more), then on further hi accumulators
    int n_lo = std::min(lo_max, bn <= 64 ? 4 : 2);
    while (n_lo > 1 && (need + n_lo) * bn > cols) --n_lo;
    *n_lo_out = n_lo;
    *n_hi_out = std::max(need, std::min(4, cols / bn - n_lo));
    return bn;
  };
  if (pair_short && n_mma_hi <= 24) {
    for (int tiles = 1; tiles <= 64; ++tiles) {
      int bn = (cout + tiles - 1) / tiles;
      bn = (bn + 15) & ~15;
      if (bn > cols_pair / 2) continue;
      return finish(bn, cols_pair, 1);
    }
  }
  // long reductions: fewest N tiles first (every extra tile is another pass over A); within a tile width, two lo
  // accumulators when the hi accumulators next to them stay under 1.25 x the MMA bound, else one
  for (int pass = 0; pass < 2; ++pass)
    for (int tiles = 1; tiles <= 64; ++tiles) {
      int bn = (cout + tiles - 1) / tiles;
      bn = (bn + 15) & ~15;
      if (bn > 256) continue;
      const int total = std::min(8, cols_full / bn);
      if (total < 2) continue;
      if (total >= 3 && lo_max >= 2) {
        const int n_hi = std::min(4, total - 2);
        if (pass == 1 || 4 * n_mma_hi <= 5 * n_hi * max_mma_per_acc()) {
          *n_hi_out = n_hi;
          *n_lo_out = std::min(lo_max, bn <= 64 ? std::min(4, total - n_hi) : 2);
          return bn;
        }
      }
      const int n_hi = std::min(4, total - 1);
      if (pass == 1 || n_mma_hi <= n_hi * max_mma_per_acc()) {
        *n_hi_out = n_hi;
        *n_lo_out = 1;
        return bn;
      }
    }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
void *pcodec_tc16_weights_create(const float *w_tap_major, int n_taps, int cin_total, int cout, void *stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || cout % 16 != 0 || cin_total % 4 != 0) return nullptr;
  W16 *h = new W16();
  h->cout = cout;
  h->k_total = n_taps * cin_total;
  h->k_stride = (h->k_total + 7) & ~7;  // TMA row pitch: a multiple of 16 bytes
  const size_t bytes = sizeof(__half) * (size_t)cout * h->k_stride;
  h->dev_hi = h->dev_lo = nullptr;
  h->dev_scale = nullptr;
  if (cudaMalloc(&h->dev_hi, bytes) != cudaSuccess || cudaMalloc(&h->dev_lo, bytes) != cudaSuccess ||
      cudaMalloc(&h->dev_scale, 2 * sizeof(float)) != cudaSuccess) {
    cudaFree(h->dev_hi);
    cudaFree(h->dev_lo);
    cudaFree(h->dev_scale);
    delete h;
    return nullptr;
  }
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(h->dev_scale, 0, 2 * sizeof(float), st);
  const int64_t n_w = (int64_t)n_taps * cin_total * cout;
  absmax_kernel<<<(unsigned)std::min<int64_t>(1024, ceil_div64(n_w, 256)), 256, 0, st>>>(w_tap_major, n_w, h->dev_scale);
  PCODEC_COUNT_LAUNCH();
  const int64_t total = (int64_t)cout * h->k_stride;
  split_weights16_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(w_tap_major, n_taps, cin_total, cout, h->k_stride,
                                                                          h->dev_scale, h->dev_hi, h->dev_lo);
  PCODEC_COUNT_LAUNCH();
  // (the TMA maps depend on the N tile and are encoded at plan time, pcodec_conv_plan)
  if (cudaGetLastError() != cudaSuccess) {
    cudaFree(h->dev_hi);
    cudaFree(h->dev_lo);
    cudaFree(h->dev_scale);
    delete h;
    return nullptr;
  }
  return h;
}

void pcodec_tc16_weights_destroy(void *handle) {
  if (!handle) return;
  W16 *h = static_cast<W16 *>(handle);
  cudaFree(h->dev_hi);
  cudaFree(h->dev_lo);
  cudaFree(h->dev_scale);
  delete h;
}

const void *pcodec_tc_w16(const void *tc_weights);  // conv_tc.cu: the fp16 weights behind a pcodec_conv_tc_prepare handle

#ifdef PCODEC_EXPERIMENTS
extern "C" int pcodec_debug_tc16_trace(long long *out, int n) {
  if (!out || n < 128) return 128;
  return cudaMemcpyFromSymbol(out, g_trace16, sizeof(long long) * 128) == cudaSuccess ? 128 : -1;
}
#endif

extern "C" int pcodec_split_planes(const float *src, int src_ps, int64_t n_pixels, int channels, uint16_t *hi, uint16_t *lo,
                                   int dst_ps, int square, void *stream) {
  if (!src || !hi || !lo || n_pixels <= 0 || channels <= 0) return PCODEC_ERR_BAD_ARG;
  if ((channels & 7) || (dst_ps & 7) || (src_ps & 3) || (reinterpret_cast<uintptr_t>(src) & 15) ||
      (reinterpret_cast<uintptr_t>(hi) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15))
    return PCODEC_ERR_UNSUPPORTED;
  const int64_t total = n_pixels * (channels >> 3);
  split_planes_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(src, src_ps, n_pixels, channels, hi, lo,
                                                                                      dst_ps, square);
  PCODEC_RETURN_LAUNCH();
}

extern "C" int pcodec_conv_plan(pcodec_conv_desc *desc) {
  if (!desc || !desc->plan) return PCODEC_ERR_BAD_ARG;
  Plan16 *pl = static_cast<Plan16 *>(desc->plan);
  pl->magic = 0;
  const W16 *w = static_cast<const W16 *>(pcodec_tc_w16(desc->tc_weights));
  EncodeTiledFn enc = encode_fn();
  if (!w || !enc) return PCODEC_ERR_UNSUPPORTED;
  if (w->cout != desc->cout || w->k_total != desc->n_taps * desc->cin_total) return PCODEC_ERR_BAD_ARG;
  if (desc->n_segments < 1 || desc->n_segments > PCODEC_MAX_SEGMENTS) return PCODEC_ERR_BAD_ARG;
  // epilogue alignment (float4 / 4 x fp16 accesses)
  const bool want_f32 = !(desc->flags & PCODEC_FLAG_NO_F32_OUT);
  if (!want_f32 && !desc->out_hi) return PCODEC_ERR_BAD_ARG;
  if (desc->flags & PCODEC_FLAG_SUBPIXEL_NCHW) {
    if (desc->cout != 16 || desc->out_pixel_stride < 1 || desc->out_pixel_stride > 4 || !desc->out ||
        (reinterpret_cast<uintptr_t>(desc->out) & 7) || (desc->out_w & 1) || desc->out_hi)
      return PCODEC_ERR_UNSUPPORTED;
  } else {
    if (want_f32 && (!desc->out || (desc->out_pixel_stride & 3) || (reinterpret_cast<uintptr_t>(desc->out) & 15)))
      return PCODEC_ERR_UNSUPPORTED;
    if (desc->out_hi && (!desc->out_lo || (desc->out_plane_stride & 3) || (reinterpret_cast<uintptr_t>(desc->out_hi) & 7) ||
                         (reinterpret_cast<uintptr_t>(desc->out_lo) & 7)))
      return PCODEC_ERR_UNSUPPORTED;
  }
  if (desc->r1 && ((desc->r1_pixel_stride & 3) || (reinterpret_cast<uintptr_t>(desc->r1) & 15))) return PCODEC_ERR_UNSUPPORTED;
  if (desc->r2 && ((desc->r2_pixel_stride & 3) || (reinterpret_cast<uintptr_t>(desc->r2) & 15))) return PCODEC_ERR_UNSUPPORTED;
  for (const pcodec_planes *rp : {&desc->r1_16, &desc->r2_16})
    if (rp->hi && (!rp->lo || (rp->pixel_stride & 3) || (reinterpret_cast<uintptr_t>(rp->hi) & 7) ||
                   (reinterpret_cast<uintptr_t>(rp->lo) & 7)))
      return PCODEC_ERR_UNSUPPORTED;
  if ((desc->r1_16.hi || desc->r2_16.hi || (desc->flags & PCODEC_FLAG_SQUARE_OUT_PLANES)) &&
      (desc->flags & (PCODEC_FLAG_PIXEL_SHUFFLE2 | PCODEC_FLAG_SUBPIXEL_NCHW)))
    return PCODEC_ERR_UNSUPPORTED;
  if (desc->in_step < 1 || desc->in_step > 2) return PCODEC_ERR_UNSUPPORTED;

  // tile shape: th x tw = 128 output-grid pixels, tw a power of two; least padded area, then the squarest
  int best_tw = 0;
  int64_t best_area = 0;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = BM / tw;
    if (tw * desc->in_step > 256 || th * desc->in_step > 256) continue;  // TMA box limits
    const int64_t area = (int64_t)((desc->grid_w + tw - 1) / tw) * tw * ((desc->grid_h + th - 1) / th) * th;
    if (!best_tw || area < best_area || (area == best_area && std::abs(tw - th) < std::abs(best_tw - BM / best_tw))) {
      best_tw = tw;
      best_area = area;
    }
  }
  if (!best_tw) return PCODEC_ERR_UNSUPPORTED;
  pl->tw = best_tw;
  pl->th = BM / best_tw;
  pl->tw_shift = 0;
  while ((1 << pl->tw_shift) < pl->tw) ++pl->tw_shift;
  pl->tiles_w = (desc->grid_w + pl->tw - 1) / pl->tw;
  pl->tiles_h = (desc->grid_h + pl->th - 1) / pl->th;
  if ((int64_t)pl->tiles_w * pl->tiles_h * desc->batch >= (1ll << 31)) return PCODEC_ERR_UNSUPPORTED;

  int n_steps = 0, n_mma_hi = 0;
  for (int s = 0; s < desc->n_segments; ++s) {
    const int ch = desc->seg[s].channels;
    if (ch <= 0 || !desc->seg16[s].hi || !desc->seg16[s].lo) return PCODEC_ERR_BAD_ARG;
    if ((desc->seg16[s].pixel_stride & 7) || (reinterpret_cast<uintptr_t>(desc->seg16[s].hi) & 15) ||
        (reinterpret_cast<uintptr_t>(desc->seg16[s].lo) & 15))
      return PCODEC_ERR_UNSUPPORTED;
    n_steps += desc->n_taps * ((ch + KS - 1) / KS);
    n_mma_hi += desc->n_taps * ((ch + 15) / 16);
  }
  pl->n_steps = n_steps;
  static const int ts_default = [] { const char *e = pcodec_knob("PCODEC_TC16_TS"); return e ? atoi(e) : 0; }();
  pl->ts = ts_default;
  pl->bn = pick_bn16(desc->cout, n_mma_hi, &pl->n_hi, &pl->n_lo, pl->ts);
  pl->n_ksteps = n_mma_hi;
  if (pl->bn == 0) return PCODEC_ERR_UNSUPPORTED;
  pl->n_tiles = (desc->cout + pl->bn - 1) / pl->bn;
  const int stage_bytes = 2 * A_BYTES + 2 * pl->bn * 128;
  auto need = [&](int st) { return std::max(st * stage_bytes, EPI_WARPS_MAX * 2048) + 8 * (2 * st + 2) + 64; };
  int tmem_cols = 32;
  while (tmem_cols < (pl->n_hi + pl->n_lo) * pl->bn + (pl->ts ? 64 : 0)) tmem_cols <<= 1;
  // tiles of <= 256 TMEM columns: keep the footprint small enough for TWO resident CTAs (one's prologue / epilogue
  // overlaps the other's main loop); wider tiles take the whole SM and as many stages as fit
  const int limit = (tmem_cols <= 256 && need(1) <= SMEM_HALF) ? SMEM_HALF : SMEM_LIMIT;
  int stages = 1;
  while (stages < 6 && need(stages + 1) <= limit) ++stages;
  if (stages > n_steps) stages = n_steps;
  if (need(stages) > SMEM_LIMIT) return PCODEC_ERR_UNSUPPORTED;
  pl->stages = stages;
  pl->smem = need(stages);
  if (tmem_cols > 256 && pl->smem < 116 * 1024) pl->smem = 116 * 1024;  // a 512-column CTA must not share the SM

  // activation planes: dims {C, W, H, N}; box {64, tw*s, th*s, 1} traversed with element strides {1, s, s, 1}
  CUtensorMap tmp;
  for (int s = 0; s < desc->n_segments; ++s) {
    const pcodec_planes &pp = desc->seg16[s];
    const cuuint64_t dims[4] = {(cuuint64_t)desc->seg[s].channels, (cuuint64_t)desc->in_w, (cuuint64_t)desc->in_h,
                                (cuuint64_t)desc->batch};
    const cuuint64_t strides[3] = {(cuuint64_t)pp.pixel_stride * 2, (cuuint64_t)desc->in_w * pp.pixel_stride * 2,
                                   (cuuint64_t)desc->in_h * desc->in_w * pp.pixel_stride * 2};
    const cuuint32_t box[4] = {(cuuint32_t)KS, (cuuint32_t)(pl->tw * desc->in_step), (cuuint32_t)(pl->th * desc->in_step), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)desc->in_step, (cuuint32_t)desc->in_step, 1};
    for (int part = 0; part < 2; ++part) {
      CUresult r = enc(&tmp, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint16_t *>(part == 0 ? pp.hi : pp.lo), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        if (getenv("PCODEC_TC_VERBOSE")) fprintf(stderr, "[conv_tc16] cuTensorMapEncodeTiled(A) failed: %d\n", (int)r);
        return PCODEC_ERR_UNSUPPORTED;
      }
      memcpy(pl->maps + sizeof(CUtensorMap) * (part * PCODEC_MAX_SEGMENTS + s), &tmp, sizeof(CUtensorMap));
    }
  }
  {
    // weight maps: box {64 K elements, bn rows}; out-of-range rows / columns are zero-filled (padded N tile, K tail)
    const cuuint64_t dims[2] = {(cuuint64_t)w->k_stride, (cuuint64_t)w->cout};
    const cuuint64_t strides[1] = {(cuuint64_t)w->k_stride * sizeof(__half)};
    const cuuint32_t box[2] = {(cuuint32_t)KS, (cuuint32_t)pl->bn};
    const cuuint32_t estr[2] = {1, 1};
    for (int part = 0; part < 2; ++part) {
      CUresult r = enc(&tmp, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, part == 0 ? (void *)w->dev_hi : (void *)w->dev_lo, dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return PCODEC_ERR_UNSUPPORTED;
      memcpy(pl->maps + sizeof(CUtensorMap) * (8 + part), &tmp, sizeof(CUtensorMap));
    }
  }
  pl->magic = PLAN_MAGIC;
  return PCODEC_OK;
}

bool pcodec_conv_taps_tc16_ready(const pcodec_conv_desc *desc) {
  return desc->plan && static_cast<const Plan16 *>(desc->plan)->magic == PLAN_MAGIC && pcodec_tc_w16(desc->tc_weights);
}

int pcodec_conv_taps_tc16(const pcodec_conv_desc *desc, void *stream) {
  if (!pcodec_conv_taps_tc16_ready(desc)) return PCODEC_ERR_UNSUPPORTED;
  const Plan16 *pl = static_cast<const Plan16 *>(desc->plan);
  const W16 *w = static_cast<const W16 *>(pcodec_tc_w16(desc->tc_weights));
  Params16 P;
  memcpy(P.a_hi, pl->maps, sizeof(CUtensorMap) * PCODEC_MAX_SEGMENTS);
  memcpy(P.a_lo, pl->maps + sizeof(CUtensorMap) * PCODEC_MAX_SEGMENTS, sizeof(CUtensorMap) * PCODEC_MAX_SEGMENTS);
  memcpy(&P.w_hi, pl->maps + sizeof(CUtensorMap) * 8, sizeof(CUtensorMap));
  memcpy(&P.w_lo, pl->maps + sizeof(CUtensorMap) * 9, sizeof(CUtensorMap));
  P.d = *desc;
  P.w_scale = w->dev_scale;
  P.tw = pl->tw; P.th = pl->th; P.tw_shift = pl->tw_shift; P.tiles_w = pl->tiles_w; P.tiles_h = pl->tiles_h;
  P.bn = pl->bn; P.stages = pl->stages; P.n_hi = pl->n_hi; P.n_lo = pl->n_lo; P.n_steps = pl->n_steps;
  P.n_ksteps = pl->n_ksteps;
  P.ts = pl->ts;
  P.debug = 0;
  if (const char *e = pcodec_knob("PCODEC_TC16_DEBUG")) P.debug = atoi(e);
  if (getenv("PCODEC_TC_VERBOSE"))
    fprintf(stderr, "[conv_tc16] grid %dx%d tile %dx%d bn=%d n_tiles=%d n_steps=%d stages=%d n_hi=%d n_lo=%d smem=%d\n", desc->grid_h,
            desc->grid_w, pl->th, pl->tw, pl->bn, pl->n_tiles, pl->n_steps, pl->stages, pl->n_hi, pl->n_lo, pl->smem);
  {
    static std::atomic<uint64_t> attr_mask{0};
    uint64_t bit;
    if (pcodec_device_needs(attr_mask, &bit)) {
      PCODEC_CHECK_CUDA(cudaFuncSetAttribute(conv_taps_tc16_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      PCODEC_CHECK_CUDA(cudaFuncSetAttribute(conv_taps_tc16_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      attr_mask.fetch_or(bit, std::memory_order_release);
    }
  }
  const int64_t n_ctas = (int64_t)pl->tiles_w * pl->tiles_h * desc->batch * pl->n_tiles;
  if (n_ctas >= (1ll << 31)) return PCODEC_ERR_UNSUPPORTED;
  dim3 grid((unsigned)n_ctas);
  // tiles that own the SM (more than 113 KB / 256 TMEM columns) get the 12-epilogue-warp instantiation
  static const bool wide_epi = [] { const char *e = pcodec_knob("PCODEC_TC16_EPI12"); return !e || atoi(e) != 0; }();
  if (wide_epi && pl->smem > SMEM_HALF)
    conv_taps_tc16_kernel<12><<<grid, 32 * 14, pl->smem, as_stream(stream)>>>(P);
  else
    conv_taps_tc16_kernel<8><<<grid, 32 * 10, pl->smem, as_stream(stream)>>>(P);
  PCODEC_RETURN_LAUNCH();
}
