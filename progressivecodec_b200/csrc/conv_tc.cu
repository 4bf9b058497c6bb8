// tcgen05 / TMEM / TMA implicit-GEMM convolution ("sum of shifted-tap GEMMs") with 3xTF32 split accumulation.
//
// Same contract as the SIMT kernel (conv_simt.cu) but the contraction runs on the 5th-gen tensor cores:
//   * CTA tile 128 (output pixels) x BN (output channels, 16..256), K slab = 32 input channels of one tap.
//   * B (weights, K-major [Cout][T*Cin] fp32, pre-split on the host into TF32 hi and lo parts) arrives by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage ring; OOB columns of the last slab are
//     zero-filled by the TMA unit.
//   * A (activations) is an on-the-fly gather of the shifted NHWC patch (padding / stride / virtual concat in the
//     address math).  Four loader warps copy it with 16-byte cp.async (LDGSTS, zero-fill for padding) into a
//     shared-memory staging tile; four converter warps read one tile ROW per thread, form lo = x - trunc_tf32(x)
//     (and x*x for GDN) and write hi (= the raw fp32: kind::tf32 ignores the low 13 mantissa bits, verified) and
//     lo with tcgen05.st into a double-buffered A operand area of TENSOR MEMORY.
//   * One elected thread issues tcgen05.mma.kind::tf32 in the TS form (A from TMEM, B from shared memory):
//     acc += Ahi*Bhi ; acc_lo += Alo*Bhi + Ahi*Blo.  A in TMEM matters: in the SS form the operand fetch from
//     shared memory (~64 B/clk) re-read the 4 KB A tile for every one of the 3 products and capped the kernel at
//     ~43 % of the MMA rate for N = 96 (measured: time per MMA == (A bytes + B bytes)/64).
//     The dropped Alo*Blo term is ~2^-22 relative, i.e. fp32-class accuracy, which the codec needs because these
//     outputs feed round(), sigma->CDF-index thresholds and the quantile ranking.
//     `tc_split = 1` issues only the first product (plain TF32) — for layers whose output only enters PSNR.
//   * tcgen05.commit releases ring slots / A buffers and finally signals the epilogue; all eight producer warps
//     then read the accumulators with tcgen05.ld (one TMEM lane = one output pixel per thread), sum the partial
//     accumulators, apply the fused epilogue (bias / GELU / residual / gate / GDN / LRP / clamp / pixel-shuffle)
//     and store 64-byte runs per thread.
// The K order (segment, tap, channel slab; hi*hi, lo*hi, hi*lo) is fixed: deterministic and batch invariant.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                  // fp32 elements per K slab = 128 bytes = one swizzle row
constexpr int TC_A_BYTES = TC_BM * 128;    // one A tile (hi or lo)
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = 32 * (TC_PRODUCER_WARPS + 2);  // warps 0-7: A producers + epilogue, warp 8: TMA + TMEM alloc, warp 9: MMA
constexpr int TC_SMEM_LIMIT = 225 * 1024;

struct TcWeights {
  CUtensorMap map_hi, map_lo;
  float *dev_hi, *dev_lo;  // [cout][k_total]
  int cout, k_total, bn, n_tiles;
};

struct TcParams {
  pcodec_conv_desc d;
  int64_t M;
  int bn, stages, split, n_steps;
  int raw_stages;  // unused (kept for layout stability)
  int n_hi_acc;  // TMEM accumulators for the hi*hi products (round-robin over K slabs); +1 for the lo terms when split
};

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 16-byte async copy global -> shared (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on an mbarrier once all prior cp.async of this thread have landed (does not bump the pending count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, M=128, N from idesc, K=8
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (TS form): A = 128 lanes x 8 columns of fp32 at tmem_a
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: this thread's lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t addr, const float4 (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(addr),
      "r"(__float_as_uint(v[0].x)), "r"(__float_as_uint(v[0].y)), "r"(__float_as_uint(v[0].z)), "r"(__float_as_uint(v[0].w)),
      "r"(__float_as_uint(v[1].x)), "r"(__float_as_uint(v[1].y)), "r"(__float_as_uint(v[1].z)), "r"(__float_as_uint(v[1].w)),
      "r"(__float_as_uint(v[2].x)), "r"(__float_as_uint(v[2].y)), "r"(__float_as_uint(v[2].z)), "r"(__float_as_uint(v[2].w)),
      "r"(__float_as_uint(v[3].x)), "r"(__float_as_uint(v[3].y)), "r"(__float_as_uint(v[3].z)), "r"(__float_as_uint(v[3].w)),
      "r"(__float_as_uint(v[4].x)), "r"(__float_as_uint(v[4].y)), "r"(__float_as_uint(v[4].z)), "r"(__float_as_uint(v[4].w)),
      "r"(__float_as_uint(v[5].x)), "r"(__float_as_uint(v[5].y)), "r"(__float_as_uint(v[5].z)), "r"(__float_as_uint(v[5].w)),
      "r"(__float_as_uint(v[6].x)), "r"(__float_as_uint(v[6].y)), "r"(__float_as_uint(v[6].z)), "r"(__float_as_uint(v[6].w)),
      "r"(__float_as_uint(v[7].x)), "r"(__float_as_uint(v[7].y)), "r"(__float_as_uint(v[7].z)), "r"(__float_as_uint(v[7].w))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// SWIZZLE_128B, K-major, 8-row x 128-byte atoms stacked along M/N with a 1024-byte stride (SM100 descriptor v1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)0 << 16;                       // leading byte offset (unused: one swizzle atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float tc_epilogue(int epi, float acc, float r1, float r2, bool has_r2) {
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
    default: return acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_taps_tc_kernel(const __grid_constant__ TcParams P, const __grid_constant__ CUtensorMap map_hi,
                    const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const pcodec_conv_desc &d = P.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, stages = P.stages;
  const bool split = P.split == 3;
  const int b_bytes = bn * 128;
  const int stage_bytes = TC_A_BYTES + (split ? 2 : 1) * b_bytes;  // raw A staging tile | B_hi | (B_lo)

  // 1024-byte aligned carve-up
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto a_raw = [&](int s) { return smem_base + s * stage_bytes; };
  auto b_hi = [&](int s) { return smem_base + s * stage_bytes + TC_A_BYTES; };
  auto b_lo = [&](int s) { return b_hi(s) + b_bytes; };
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };                   // cp.async landed        (128 loaders)
  auto raw_empty = [&](int s) { return bar_base + 8u * (stages + s); };       // staging tile consumed  (128 converters)
  auto full_b = [&](int s) { return bar_base + 8u * (2 * stages + s); };      // TMA bytes landed
  auto empty_b = [&](int s) { return bar_base + 8u * (3 * stages + s); };     // MMAs that read B done  (tcgen05.commit)
  auto a_full = [&](int q) { return bar_base + 8u * (4 * stages + q); };      // A operand in TMEM      (128 converters)
  auto a_empty = [&](int q) { return bar_base + 8u * (4 * stages + 2 + q); }; // MMAs that read it done (tcgen05.commit)
  const uint32_t tmem_full = bar_base + 8u * (4 * stages + 4);
  const uint32_t tmem_slot = tmem_full + 8u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base

  // TMEM layout: [n_acc accumulators of bn columns][2 A-operand buffers of a_cols columns].
  // Accumulators: the tensor core's fp32 accumulate truncates (round-toward-zero) once per MMA, so the error grows
  // linearly with the number of MMAs that touch an accumulator.  The small lo*hi / hi*lo products therefore get
  // their own accumulator (their truncation error is 2^-11 smaller in absolute terms), and the hi*hi products
  // round-robin over n_hi_acc accumulators; the epilogue sums them with round-to-nearest adds.
  const int n_acc = P.n_hi_acc + (split ? 1 : 0);
  const int a_cols = split ? 64 : 32;            // hi (32 fp32 columns) + lo (32)
  const uint32_t a_tmem_off = (uint32_t)(n_acc * bn);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_acc * bn + 2 * a_cols) tmem_cols <<= 1;

  __shared__ int s_dy[PCODEC_MAX_TAPS], s_dx[PCODEC_MAX_TAPS];
  __shared__ const float *s_seg_ptr[PCODEC_MAX_SEGMENTS];
  __shared__ int s_seg_ch[PCODEC_MAX_SEGMENTS], s_seg_ps[PCODEC_MAX_SEGMENTS];
  if (threadIdx.x < PCODEC_MAX_TAPS) {
    s_dy[threadIdx.x] = threadIdx.x < d.n_taps ? d.dy[threadIdx.x] : 0;
    s_dx[threadIdx.x] = threadIdx.x < d.n_taps ? d.dx[threadIdx.x] : 0;
  }
  if (threadIdx.x < PCODEC_MAX_SEGMENTS) {
    const bool in = (int)threadIdx.x < d.n_segments;
    s_seg_ptr[threadIdx.x] = in ? d.seg[threadIdx.x].ptr : nullptr;
    s_seg_ch[threadIdx.x] = in ? d.seg[threadIdx.x].channels : 0;
    s_seg_ps[threadIdx.x] = in ? d.seg[threadIdx.x].pixel_stride : 0;
  }

  if (warp == TC_PRODUCER_WARPS + 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(raw_full(s), 128);
      mbar_init(raw_empty(s), 128);
      mbar_init(full_b(s), 1);
      mbar_init(empty_b(s), 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(a_full(q), 128);
      mbar_init(a_empty(q), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == TC_PRODUCER_WARPS) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * bn;
  const int n_steps = P.n_steps;

  if (warp < TC_PRODUCER_WARPS) {
    // =============================== A producers ===============================
    if (warp < 4) {
      // ------------------------------- loaders (warps 0-3) -------------------------------
      // 16-byte cp.async of the shifted patch into the slot's staging tile, 128-byte-swizzled ([row][chunk ^ row&7])
      // so that the converters' row-wise reads are bank-conflict free.  A lane owns chunk (lane & 7) of 8 rows.
      constexpr int RPT = 8;
      const int chunk = lane & 7, sub = lane >> 3;
      uint32_t soff[RPT];
      int pix0[RPT], ih0[RPT], iw0[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = warp * 32 + i * 4 + sub;
        soff[i] = r * 128 + ((chunk ^ (r & 7)) << 4);
        const int64_t m = m0 + r;
        const bool okr = m < P.M;
        const int64_t mm = okr ? m : 0;
        const int w = (int)(mm % d.grid_w);
        const int64_t t = mm / d.grid_w;
        const int h = (int)(t % d.grid_h);
        const int n = (int)(t / d.grid_h);
        ih0[i] = okr ? h * d.in_step : -(1 << 28);
        iw0[i] = w * d.in_step;
        pix0[i] = (n * d.in_h + h * d.in_step) * d.in_w + w * d.in_step;
      }
      const int in_h = d.in_h, in_w = d.in_w, n_taps = d.n_taps;
      int seg = 0, tap = 0, st = 0;
      uint32_t eph = 1;  // parity to wait for on raw_empty
      for (int s = 0; s < n_steps;) {
        const float *base = s_seg_ptr[seg] + chunk * 4;
        const int ps = s_seg_ps[seg], seg_channels = s_seg_ch[seg];
        const int seg_slabs = (seg_channels + TC_BK - 1) / TC_BK;
        const int dy = s_dy[tap], dx = s_dx[tap];
        const int tapoff = dy * in_w + dx;
        const float *rowptr[RPT];
        uint32_t rowmask = 0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const bool v = (unsigned)(ih0[i] + dy) < (unsigned)in_h && (unsigned)(iw0[i] + dx) < (unsigned)in_w;
          rowptr[i] = base + (int64_t)(v ? pix0[i] + tapoff : 0) * ps;
          rowmask |= (v ? 1u : 0u) << i;
        }
        for (int kc = 0; kc < seg_slabs; ++kc, ++s) {
          mbar_wait(raw_empty(st), eph);
          const uint32_t dst = a_raw(st);
          const uint32_t mask = (kc * TC_BK + chunk * 4 < seg_channels) ? rowmask : 0u;
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            cp_async16(dst + soff[i], rowptr[i] + kc * TC_BK, ((mask >> i) & 1u) ? 16u : 0u);
          cp_async_arrive_noinc(raw_full(st));
          if (++st == stages) { st = 0; eph ^= 1u; }
        }
        if (++tap == n_taps) { tap = 0; ++seg; }
      }
    } else {
      // ------------------------------- converters (warps 4-7) -------------------------------
      // Thread = tile row (TMEM lane) 32*(warp & 3) + lane: read the row's 32 floats from the staging tile, write
      // hi (raw; the MMA truncates) and lo = x - trunc_tf32(x) into the A operand buffer of tensor memory.
      const bool square = (d.flags & PCODEC_FLAG_SQUARE_INPUT) != 0;
      const int arow = (warp & 3) * 32 + lane;
      const uint32_t row_off = arow * 128, row_x = (uint32_t)(arow & 7);
      const uint32_t lane_base = tmem_acc + a_tmem_off + ((uint32_t)((warp & 3) * 32) << 16);
      int st = 0, q = 0;
      uint32_t ph = 0, qph = 1;
      for (int s = 0; s < n_steps; ++s) {
        mbar_wait(raw_full(st), ph);
        const uint8_t *src = smem_gen + (size_t)st * stage_bytes + row_off;
        float4 x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4 *>(src + (((uint32_t)c ^ row_x) << 4));
        mbar_arrive(raw_empty(st));  // staging tile consumed (values are in registers)
        if (square) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { x[c].x *= x[c].x; x[c].y *= x[c].y; x[c].z *= x[c].z; x[c].w *= x[c].w; }
        }
        mbar_wait(a_empty(q), qph);  // MMAs that read this A buffer two slabs ago have retired
        tc_fence_after();
        const uint32_t ta = lane_base + (uint32_t)(q * a_cols);
        tmem_st32(ta, x);
        if (split) {
          float4 l[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            l[c].x = x[c].x - __uint_as_float(__float_as_uint(x[c].x) & 0xFFFFE000u);
            l[c].y = x[c].y - __uint_as_float(__float_as_uint(x[c].y) & 0xFFFFE000u);
            l[c].z = x[c].z - __uint_as_float(__float_as_uint(x[c].z) & 0xFFFFE000u);
            l[c].w = x[c].w - __uint_as_float(__float_as_uint(x[c].w) & 0xFFFFE000u);
          }
          tmem_st32(ta + 32u, l);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(a_full(q));
        if (++st == stages) { st = 0; ph ^= 1u; }
        q ^= 1;
        if (q == 0) qph ^= 1u;
      }
    }

    // =============================== epilogue ===============================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;      // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int half = warp >> 2;        // warps w and w+4 share a quarter and interleave the 16-column groups
    const int row = quarter * 32 + lane;  // TMEM lane == tile row
    const int64_t m = m0 + row;
    const bool row_ok = m < P.M;
    const int64_t mm = row_ok ? m : 0;
    const int w = (int)(mm % d.grid_w);
    const int64_t t = mm / d.grid_w;
    const int h = (int)(t % d.grid_h);
    const int64_t n = t / d.grid_h;
    const int oh = h * d.out_step + d.out_off_y, ow = w * d.out_step + d.out_off_x;
    const int64_t opix = (n * d.out_h + oh) * (int64_t)d.out_w + ow;
    const bool shuffle = (d.flags & PCODEC_FLAG_PIXEL_SHUFFLE2) != 0;
    const bool has_r2 = d.r2 != nullptr;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = half * 16; c0 < bn; c0 += 32) {
      float acc[16];
      tmem_ld16(lane_addr + (uint32_t)c0, acc);  // warp-collective: executed by all lanes, stores are predicated
      for (int a = 1; a < n_acc; ++a) {
        float part[16];
        tmem_ld16(lane_addr + (uint32_t)(a * bn + c0), part);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += part[j];
      }
      if (!row_ok) continue;
      const int co0 = n0 + c0;
      float bias[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bias[j] = d.bias ? __ldg(d.bias + co0 + j) : 0.f;
      if (!shuffle) {
        float r1[16], r2[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 a = d.r1 ? *reinterpret_cast<const float4 *>(d.r1 + opix * d.r1_pixel_stride + co0 + 4 * q)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 b = d.r2 ? *reinterpret_cast<const float4 *>(d.r2 + opix * d.r2_pixel_stride + co0 + 4 * q)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          r1[4 * q] = a.x; r1[4 * q + 1] = a.y; r1[4 * q + 2] = a.z; r1[4 * q + 3] = a.w;
          r2[4 * q] = b.x; r2[4 * q + 1] = b.y; r2[4 * q + 2] = b.z; r2[4 * q + 3] = b.w;
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = tc_epilogue(d.epilogue, acc[j] + bias[j], r1[j], r2[j], has_r2);
        float *dst = d.out + opix * d.out_pixel_stride + co0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = co0 + j;
          const float v = tc_epilogue(d.epilogue, acc[j] + bias[j], 0.f, 0.f, false);
          const int c = co >> 2, si = (co >> 1) & 1, sj = co & 1;
          const int64_t sp = (n * d.out_h + (2 * oh + si)) * (int64_t)d.out_w + (2 * ow + sj);
          d.out[sp * d.out_pixel_stride + c] = v;
        }
      }
    }
    tc_fence_before();
  } else if (warp == TC_PRODUCER_WARPS) {
    // =============================== TMA producer for B ===============================
    if (lane == 0) {
      int seg = 0, tap = 0, kc = 0, seg_cbase = 0;
      int st = 0;
      uint32_t ph = 1;
      for (int s = 0; s < n_steps; ++s, ++st) {
        if (st == stages) { st = 0; ph ^= 1u; }
        mbar_wait(empty_b(st), ph);
        const int k = tap * d.cin_total + seg_cbase + kc * TC_BK;
        mbar_expect_tx(full_b(st), (uint32_t)((split ? 2 : 1) * b_bytes));
        tma_load_2d(b_hi(st), &map_hi, full_b(st), k, n0);
        if (split) tma_load_2d(b_lo(st), &map_lo, full_b(st), k, n0);
        const int sc = d.seg[seg].channels;
        if (++kc == (sc + TC_BK - 1) / TC_BK) {
          kc = 0;
          if (++tap == d.n_taps) { tap = 0; seg_cbase += sc; ++seg; }
        }
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int st = 0, hi_idx = 0, q = 0;
      uint32_t ph = 0, qph = 0;
      const uint32_t acc_lo = tmem_acc + (uint32_t)(P.n_hi_acc * bn);
      for (int s = 0; s < n_steps; ++s) {
        mbar_wait(a_full(q), qph);
        mbar_wait(full_b(st), ph);
        tc_fence_after();
        const uint64_t db_hi = umma_desc_sw128(b_hi(st)), db_lo = umma_desc_sw128(b_lo(st));
        const uint32_t ta_hi = tmem_acc + a_tmem_off + (uint32_t)(q * a_cols);
        const uint32_t ta_lo = ta_hi + 32u;
        const uint32_t acc_hi = tmem_acc + (uint32_t)(hi_idx * bn);
        const bool first_hi = s < P.n_hi_acc;  // first slab that touches this hi accumulator
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);  // B: 8 tf32 = 32 bytes = 2 x 16-byte units inside the swizzle row
          const uint32_t ak = (uint32_t)(k * 8);   // A: 8 fp32 columns of tensor memory
          umma_tf32_ts(acc_hi, ta_hi + ak, db_hi + adv, idesc, (!first_hi || k > 0) ? 1u : 0u);
          if (split) {
            umma_tf32_ts(acc_lo, ta_lo + ak, db_hi + adv, idesc, (s > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts(acc_lo, ta_hi + ak, db_lo + adv, idesc, 1u);
          }
        }
        umma_commit(empty_b(st));  // B slot reusable once these MMAs retire (implies tcgen05.fence::before_thread_sync)
        umma_commit(a_empty(q));   // A operand buffer reusable
        if (++st == stages) { st = 0; ph ^= 1u; }
        if (++hi_idx == P.n_hi_acc) hi_idx = 0;
        q ^= 1;
        if (q == 0) qph ^= 1u;
      }
      umma_commit(tmem_full);
    }
  }
  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

__global__ void split_weights_kernel(const float *__restrict__ w_tap_major, int n_taps, int cin, int cout,
                                     float *__restrict__ hi, float *__restrict__ lo) {
  // in: [tap][cin][cout]; out: [cout][tap*cin + ci] as TF32 hi (round to nearest even) and lo = w - hi
  const int64_t total = (int64_t)n_taps * cin * cout;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % ((int64_t)n_taps * cin));
  const int co = (int)(i / ((int64_t)n_taps * cin));
  const float w = w_tap_major[(int64_t)k * cout + co];
  uint32_t u = __float_as_uint(w);
  u += 0x00000FFFu + ((u >> 13) & 1u);  // round to nearest even at bit 13
  u &= 0xFFFFE000u;
  const float h = __uint_as_float(u);
  hi[i] = h;
  lo[i] = w - h;
}

// N tile: the widest multiple of 16 that divides cout and fits the TMEM budget.  Long reductions (K >= 1024) get
// a narrower tile (<= 128 columns) so that 3 hi accumulators + 1 lo accumulator fit in the 512 TMEM columns: the
// tensor core truncates once per MMA per accumulator, so spreading the K slabs over more accumulators keeps the
// result at fp32-class accuracy (measured: rms 1e-5 -> 3e-6 at K = 4800).
int pick_bn(int cout, int k_total) {
  if (cout % 16 != 0) return 0;
  int cap = k_total >= 1024 ? 128 : 192;  // 2 accumulators x 192 + 128 A-operand columns = 512 TMEM columns
  if (const char *e = getenv("PCODEC_TC_BNCAP")) cap = atoi(e);  // experiment knob
  int fallback = 0;
  for (int tiles = 1; tiles <= 16; ++tiles) {
    if (cout % tiles) continue;
    const int bn = cout / tiles;
    if (bn % 16 != 0) continue;
    if (bn <= 256 && fallback == 0) fallback = bn;
    if (bn <= cap) return (fallback != 0 && tiles > 2 * (cout / fallback)) ? fallback : bn;
  }
  return fallback;
}

}  // namespace

extern "C" int pcodec_conv_tc_prepare(const float *w_tap_major, int n_taps, int cin_total, int cout, void **handle_out,
                                      void *stream) {
  if (!w_tap_major || !handle_out || n_taps < 1 || cin_total < 4 || cout < 1) return PCODEC_ERR_BAD_ARG;
  *handle_out = nullptr;
  const int bn = pick_bn(cout, n_taps * cin_total);
  if (bn == 0 || (cin_total % 4) != 0) return PCODEC_ERR_UNSUPPORTED;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PCODEC_ERR_UNSUPPORTED;
  TcWeights *h = new TcWeights();
  h->cout = cout;
  h->k_total = n_taps * cin_total;
  h->bn = bn;
  h->n_tiles = cout / bn;
  const size_t bytes = sizeof(float) * (size_t)cout * h->k_total;
  if (cudaMalloc(&h->dev_hi, bytes) != cudaSuccess || cudaMalloc(&h->dev_lo, bytes) != cudaSuccess) {
    delete h;
    return -(int)cudaErrorMemoryAllocation;
  }
  const int64_t total = (int64_t)cout * h->k_total;
  split_weights_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(w_tap_major, n_taps, cin_total,
                                                                                       cout, h->dev_hi, h->dev_lo);
  PCODEC_COUNT_LAUNCH();
  const cuuint64_t dims[2] = {(cuuint64_t)h->k_total, (cuuint64_t)cout};
  const cuuint64_t strides[1] = {(cuuint64_t)h->k_total * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)bn};
  const cuuint32_t estr[2] = {1, 1};
  for (int part = 0; part < 2; ++part) {
    CUresult r = enc(part == 0 ? &h->map_hi : &h->map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     part == 0 ? (void *)h->dev_hi : (void *)h->dev_lo, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      cudaFree(h->dev_hi);
      cudaFree(h->dev_lo);
      delete h;
      return PCODEC_ERR_UNSUPPORTED;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int)e;
  *handle_out = h;
  return PCODEC_OK;
}

extern "C" void pcodec_conv_tc_release(void *handle) {
  if (!handle) return;
  TcWeights *h = static_cast<TcWeights *>(handle);
  cudaFree(h->dev_hi);
  cudaFree(h->dev_lo);
  delete h;
}

bool pcodec_conv_taps_tc_supported(const pcodec_conv_desc *d) {
  if (!d->tc_weights) return false;
  const TcWeights *h = static_cast<const TcWeights *>(d->tc_weights);
  if (h->cout != d->cout || h->k_total != d->n_taps * d->cin_total) return false;
  if (d->tc_split != 1 && d->tc_split != 3) return false;
  // float4 epilogue accesses
  if ((d->out_pixel_stride & 3) || (reinterpret_cast<uintptr_t>(d->out) & 15)) return false;
  if (d->r1 && ((d->r1_pixel_stride & 3) || (reinterpret_cast<uintptr_t>(d->r1) & 15))) return false;
  if (d->r2 && ((d->r2_pixel_stride & 3) || (reinterpret_cast<uintptr_t>(d->r2) & 15))) return false;
  return true;
}

int pcodec_conv_taps_tc(const pcodec_conv_desc *desc, void *stream) {
  const TcWeights *h = static_cast<const TcWeights *>(desc->tc_weights);
  TcParams P;
  P.d = *desc;
  P.M = (int64_t)desc->batch * desc->grid_h * desc->grid_w;
  P.bn = h->bn;
  P.split = desc->tc_split;
  int n_steps = 0;
  for (int s = 0; s < desc->n_segments; ++s) n_steps += desc->n_taps * ((desc->seg[s].channels + TC_BK - 1) / TC_BK);
  P.n_steps = n_steps;
  const bool split3 = P.split == 3;
  const int stage_bytes = TC_A_BYTES + (split3 ? 2 : 1) * h->bn * 128;  // raw A staging tile | B_hi | (B_lo)
  auto need = [&](int st) { return st * stage_bytes + 1024 + 8 * (4 * st + 6) + 64; };
  int stages = 2;
  while (need(stages + 1) <= TC_SMEM_LIMIT && stages < 8) ++stages;
  if (const char *e = getenv("PCODEC_TC_STAGES")) stages = std::min(stages, atoi(e));  // experiment knob
  if (stages > n_steps) stages = n_steps;
  if (stages < 1 || need(stages) > TC_SMEM_LIMIT) return PCODEC_ERR_UNSUPPORTED;
  P.stages = stages;
  P.raw_stages = 0;
  {
    // TMEM columns: n_acc accumulators of bn + two A operand buffers (hi 32 [+ lo 32] columns each)
    const int a_cols = split3 ? 64 : 32;
    int n_acc = (512 - 2 * a_cols) / h->bn;
    int n_hi = n_acc - (split3 ? 1 : 0);
    if (n_hi > 4) n_hi = 4;
    if (n_hi > n_steps) n_hi = n_steps;
    if (n_hi < 1) return PCODEC_ERR_UNSUPPORTED;
    P.n_hi_acc = n_hi;
  }
  const int smem = need(stages);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_taps_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
  });
  if (attr_err != cudaSuccess) return -(int)attr_err;
  dim3 grid((unsigned)ceil_div64(P.M, TC_BM), (unsigned)h->n_tiles);
  conv_taps_tc_kernel<<<grid, TC_THREADS, smem, as_stream(stream)>>>(P, h->map_hi, h->map_lo);
  PCODEC_RETURN_LAUNCH();
}
