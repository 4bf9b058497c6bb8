// tcgen05 / TMEM / TMA implicit-GEMM convolution ("sum of shifted-tap GEMMs") with 3xTF32 split accumulation.
//
// Same contract as the SIMT kernel (conv_simt.cu) but the contraction runs on the 5th-gen tensor cores:
//   * CTA tile 128 (output pixels) x BN (output channels, 16..256, chosen per layer so that enough accumulators fit
//     tensor memory; the last N tile may be padded), K slab = 32 input channels of one tap.
//   * B (weights, K-major [Cout][T*Cin] fp32, pre-split on the host into TF32 hi and lo parts) arrives by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage ring; out-of-range rows / columns are zero-filled
//     by the TMA unit.
//   * A (activations) is an on-the-fly gather of the shifted NHWC patch (padding / stride / virtual concat in the
//     address math).  Four loader warps copy it with 16-byte cp.async (LDGSTS, zero-fill for padding) into a
//     shared-memory staging tile; NG groups of four converter warps take alternate K slabs, read one tile ROW per
//     thread, form lo = x - trunc_tf32(x) (and x*x for GDN) and write hi (= the raw fp32: kind::tf32 ignores the low
//     13 mantissa bits, verified) and lo with tcgen05.st into a 2..3-deep A operand ring in TENSOR MEMORY.
//   * One thread chosen with elect.sync runs the whole MMA issue loop: tcgen05.mma.kind::tf32 in the TS form (A from
//     TMEM, B from shared memory), acc_hi[s % n_hi] += Ahi*Bhi ; acc_lo += Alo*Bhi + Ahi*Blo, 12 instructions per slab
//     issued straight-line, then ONE tcgen05.commit that releases both the weight stage and the A buffer.
//     A in TMEM matters: in the SS form the operand fetch from shared memory re-read the 4 KB A tile for each of the 3
//     products.  Several accumulators matter: the tensor core truncates (RZ) once per MMA per accumulator, so long
//     reductions are spread over up to 4 hi accumulators + 1 lo accumulator and summed with RN adds in the epilogue.
//     The dropped Alo*Blo term is ~2^-22 relative, i.e. fp32-class accuracy, which the codec needs because these
//     outputs feed round(), sigma->CDF-index thresholds and the quantile ranking.
//     `tc_split = 1` issues only the first product (plain TF32).
//   * Epilogue by all producer warps: tcgen05.ld (one TMEM lane = one output pixel per thread), sum of the partial
//     accumulators, a 32x16 transpose through shared memory so that residual loads / output stores are 64-byte runs,
//     fused bias / GELU / LeakyReLU / residual / gate / GDN / LRP / clamp (one vectorised switch), or the
//     pixel-shuffle / sub-pixel-NCHW stores.  Residual tiles are prefetched into L2 at kernel start.
//   * Two instantiations: <2> 14 warps, one CTA per SM (long reductions); <1> 10 warps, <= 256 TMEM columns, two CTAs
//     per SM (1-tap short reductions whose output fits one tile; see pcodec_conv_tc_prepare).
// The K order (segment, tap, channel slab; hi*hi, lo*hi, hi*lo) is fixed: deterministic and batch invariant.
// What shaped this design is recorded in DESIGN.md §3.1 (measurements: tools/trace_tc.py, tools/ubench/mma_rate.cu).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

void *pcodec_tc16_weights_create(const float *w_tap_major, int n_taps, int cin_total, int cout, void *stream);  // conv_tc16.cu
void pcodec_tc16_weights_destroy(void *handle);

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                  // fp32 elements per K slab = 128 bytes = one swizzle row
constexpr int TC_A_BYTES = TC_BM * 128;    // one A tile (hi or lo)
// warps 0..4+4*NG-1: A producers (4 loaders + NG converter groups) + epilogue; next warp: TMA + TMEM alloc; last: MMA issuer
constexpr int TC_SMEM_LIMIT = 225 * 1024;

struct TcWeights {
  CUtensorMap map_hi, map_lo;
  float *dev_hi, *dev_lo;  // [cout][k_total]
  int cout, k_total, bn, n_tiles;
  void *w16;  // fp16-split copy of the same weights for conv_tc16.cu (nullptr when that kernel cannot take them)
  int small;  // two-CTAs-per-SM variant: 1 = separate lo accumulator (bn <= 64), 2 = lo products share the hi accumulator (bn <= 128)
};

struct TcParams {
  pcodec_conv_desc d;
  int64_t M;
  int bn, stages, split, n_steps;
  int debug;  // PCODEC_TC_DEBUG bits: 1 = skip A global loads, 2 = skip B TMA loads, 4 = skip converter TMEM stores (timing
              // experiments, wrong results), 64 = record the clock64 timeline of one CTA (128: a mid-grid CTA)
  int n_hi_acc;  // TMEM accumulators for the hi*hi products (round-robin over K slabs); +1 for the lo terms when split
  int a_ring;    // depth of the A operand ring in tensor memory (2..4)
  int shared_lo;  // 1: the lo products accumulate into hi accumulator 0 (short reductions: <= ~300 MMAs in total)
  uint32_t magic_w, magic_h;  // division by grid_w / grid_h as one multiply (dividends < 2^31): q = (m * magic) >> (31 + shift)
  int shift_w, shift_h;
};

using namespace pcodec_tc;

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
// Debug timeline (PCODEC_TC_DEBUG bit 6): CTA (0,0) records clock64() at pipeline events of its first 128 slabs.
constexpr int TC_TRACE_SLABS = 128, TC_TRACE_EVENTS = 12;
__device__ long long g_tc_trace[TC_TRACE_SLABS * TC_TRACE_EVENTS + 32];
#define TC_TRACE(slab, ev)                                                                        \
  do {                                                                                            \
    if (trace && (slab) < TC_TRACE_SLABS) g_tc_trace[(slab) * TC_TRACE_EVENTS + (ev)] = clock64(); \
  } while (0)
#define TC_TRACE_G(ev)                                                                \
  do {                                                                                \
    if (trace) g_tc_trace[TC_TRACE_SLABS * TC_TRACE_EVENTS + (ev)] = clock64();        \
  } while (0)
// NG = number of converter groups.  NG = 2 (14 warps, one CTA per SM) is the throughput configuration for long
// reductions.  NG = 1 (10 warps, <= 102 registers) lets TWO CTAs share an SM when the tile needs <= ~110 KB of
// shared memory and <= 256 TMEM columns: short reductions spend most of their time in the (serial) prologue and
// epilogue of a tile, which a second resident CTA overlaps with its own main loop.
template <int NG>
__global__ void __launch_bounds__(32 * (4 + 4 * NG + 2), NG == 1 ? 2 : 1)
conv_taps_tc_kernel(const __grid_constant__ TcParams P, const __grid_constant__ CUtensorMap map_hi,
                    const __grid_constant__ CUtensorMap map_lo) {
  constexpr int TC_PRODUCER_WARPS = 4 + 4 * NG;  // 4 loader warps + NG groups of 4 converter warps
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const pcodec_conv_desc &d = P.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, stages = P.stages;
  const bool trace = (P.debug & 64) && blockIdx.x == ((P.debug & 128) ? gridDim.x / 2 : 0) &&
                     blockIdx.y == ((P.debug & 128) ? gridDim.y - 1 : 0) && lane == 0;
  if (threadIdx.x == 0) TC_TRACE_G(0);
  const bool split = P.split == 3;
  const int b_bytes = bn * 128;
  const int stage_bytes = TC_A_BYTES + (split ? 2 : 1) * b_bytes;  // raw A staging tile | B_hi | (B_lo)

  // 1024-byte aligned carve-up
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto a_raw = [&](int s) { return smem_base + s * stage_bytes; };
  auto b_hi = [&](int s) { return smem_base + s * stage_bytes + TC_A_BYTES; };
  auto b_lo = [&](int s) { return b_hi(s) + b_bytes; };
  // the stage area doubles as the epilogue's transpose scratch (2 KB per producer warp)
  const uint32_t bar_base = smem_base + max(stages * stage_bytes, TC_PRODUCER_WARPS * 2048);
  auto raw_full = [&](int s) { return bar_base + 8u * s; };                   // cp.async landed        (128 loader threads)
  auto raw_empty = [&](int s) { return bar_base + 8u * (stages + s); };       // staging tile consumed  (4 converter warps)
  auto full_b = [&](int s) { return bar_base + 8u * (2 * stages + s); };      // TMA bytes landed
  // `empty_b(st)` = "slab in ring slot st consumed": ONE tcgen05.commit per slab releases both the B stage (to the
  // TMA producer, a ring of `stages`) and the A operand buffer (to the converter group, which waits for the slot of
  // slab s - a_ring).  A commit costs the issuing thread ~150 clk, so one per slab instead of two matters.
  auto empty_b = [&](int s) { return bar_base + 8u * (3 * stages + s); };     // MMAs of the slab in this slot done (tcgen05.commit)
  auto a_full = [&](int q) { return bar_base + 8u * (4 * stages + q); };      // A ring slot q (0..3) filled (4 converter warps)
  const uint32_t tmem_full = bar_base + 8u * (4 * stages + 4);
  const uint32_t tmem_slot = tmem_full + 8u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base

  // TMEM layout: [n_acc accumulators of bn columns][a_ring A-operand buffers of a_cols columns].
  // Accumulators: the tensor core's fp32 accumulate truncates (round-toward-zero) once per MMA, so the error grows
  // linearly with the number of MMAs that touch an accumulator.  The small lo*hi / hi*lo products therefore get
  // their own accumulator (their truncation error is 2^-11 smaller in absolute terms), and the hi*hi products
  // round-robin over n_hi_acc accumulators; the epilogue sums them with round-to-nearest adds.
  // A ring: refilling an A buffer after its MMAs retire takes ~800 clk (commit -> converter wake-up -> tcgen05.st ->
  // wait::st -> arrive -> issuer wake-up) against 480..670 clk of MMA work per slab, so two buffers leave the tensor
  // pipe idle half the time (measured with the clock64 timeline, tools/trace_tc.py); three hide the round trip.
  const int n_acc = P.n_hi_acc + ((split && !P.shared_lo) ? 1 : 0);
  const int a_cols = split ? 64 : 32;            // hi (32 fp32 columns) + lo (32)
  const int a_ring = P.a_ring;
  const uint32_t a_tmem_off = (uint32_t)(n_acc * bn);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_acc * bn + a_ring * a_cols) tmem_cols <<= 1;

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
      mbar_init(raw_full(s), 128); // cp.async completion arrives, one per loader thread
      mbar_init(raw_empty(s), 4);  // one arrival per converter warp
      mbar_init(full_b(s), 1);
      mbar_init(empty_b(s), 1);  // one tcgen05.commit (by the issuer warp that owns the slab)
    }
    for (int q = 0; q < 4; ++q) mbar_init(a_full(q), 4);  // a_full(0..3) (the a_empty slots are reused)
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == TC_PRODUCER_WARPS) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
  if (threadIdx.x == 0) TC_TRACE_G(1);

  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * bn;
  const int n_steps = P.n_steps;

  if (warp < TC_PRODUCER_WARPS && (d.r1 || d.r2) && !(d.flags & PCODEC_FLAG_PIXEL_SHUFFLE2)) {
    // The epilogue reads 128 x bn residual values that do not depend on the accumulators: pull them into L2 now so
    // the epilogue's loads are L2 hits instead of a chain of serialised DRAM round trips (one CTA per SM: nothing
    // else hides that latency).
    const int prow = threadIdx.x & 127, psub = threadIdx.x >> 7;  // 3 threads per tile row
    const int64_t pm = m0 + prow;
    if (pm < P.M) {
      const uint32_t pt = fast_div((uint32_t)pm, P.magic_w, P.shift_w);
      const int pw = (int)((uint32_t)pm - pt * (uint32_t)d.grid_w);
      const uint32_t pnn = fast_div(pt, P.magic_h, P.shift_h);
      const int ph_ = (int)(pt - pnn * (uint32_t)d.grid_h);
      const int64_t pn = (int64_t)pnn;
      const int64_t ppix = (pn * d.out_h + (ph_ * d.out_step + d.out_off_y)) * (int64_t)d.out_w + (pw * d.out_step + d.out_off_x);
      for (int c = psub * 32; c < bn && n0 + c < d.cout; c += 32 * (TC_PRODUCER_WARPS / 4)) {
        if (d.r1) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r1 + ppix * d.r1_pixel_stride + n0 + c));
        if (d.r2) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.r2 + ppix * d.r2_pixel_stride + n0 + c));
      }
    }
  }
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
        const uint32_t mm = okr ? (uint32_t)m : 0u;  // M < 2^31 (checked on the host): 32-bit divisions
        const uint32_t t = fast_div(mm, P.magic_w, P.shift_w);
        const int w = (int)(mm - t * (uint32_t)d.grid_w);
        const int n = (int)fast_div(t, P.magic_h, P.shift_h);
        const int h = (int)(t - (uint32_t)n * (uint32_t)d.grid_h);
        ih0[i] = okr ? h * d.in_step : -(1 << 28);
        iw0[i] = w * d.in_step;
        pix0[i] = (n * d.in_h + h * d.in_step) * d.in_w + w * d.in_step;
      }
      const int in_h = d.in_h, in_w = d.in_w, n_taps = d.n_taps;
      int seg = 0, tap = 0, st = 0;
      uint32_t eph = 1;  // parity to wait for on raw_empty
      for (int s = 0; s < n_steps;) {
        const float *seg_base = s_seg_ptr[seg];
        const float *base = seg_base + chunk * 4;
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
          mbar_wait_warp(raw_empty(st), eph, lane);
          if (warp == 0) TC_TRACE(s, 0);
          const uint32_t dst = a_raw(st);
          const uint32_t mask = (kc * TC_BK + chunk * 4 < seg_channels) ? rowmask : 0u;
          if (!(P.debug & 1)) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              // zero-fill copies (padding rows, channel chunks past the segment) never form an address outside the
              // segment: a 0-byte LDGSTS must not depend on what the hardware does with an unmapped source address
              const bool okc = (mask >> i) & 1u;
              cp_async16(dst + soff[i], okc ? rowptr[i] + kc * TC_BK : seg_base, okc ? 16u : 0u);
            }
          }
          cp_async_arrive_noinc(raw_full(st));  // per-thread: fires when this thread's copies have landed
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
      // Two converter groups (warps 4-7 / 8-11) take alternate K slabs, each with its own A operand buffer: one
      // group's wait -> LDS -> tcgen05.st -> wait::st -> arrive chain (latency bound, ~1 us) overlaps the other's.
      const int grp = (warp - 4) >> 2;
      for (int s = grp; s < n_steps; s += NG) {
        const int q = s % a_ring;
        const int st = s % stages;
        const uint32_t ph = (uint32_t)(s / stages) & 1u;
        if ((warp & 3) == 0) TC_TRACE(s, 2);
        {  // lane 0: A staging tile landed; lane 1: this slab's weights landed (so a_full tells the MMA warps both)
          const uint32_t wbar = lane == 0 ? raw_full(st) : full_b(st);
          if (lane < 2) mbar_wait(wbar, ph);
          __syncwarp();
        }
        if ((warp & 3) == 0) TC_TRACE(s, 3);
        const uint8_t *src = smem_gen + (size_t)st * stage_bytes + row_off;
        float4 x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4 *>(src + (((uint32_t)c ^ row_x) << 4));
        if (square) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { x[c].x *= x[c].x; x[c].y *= x[c].y; x[c].z *= x[c].z; x[c].w *= x[c].w; }
        }
        if (s >= a_ring) {  // MMAs that read this A buffer a_ring slabs ago have retired
          const int sp = s - a_ring;
          mbar_wait_warp(empty_b(sp % stages), (uint32_t)(sp / stages) & 1u, lane);
        }
        if ((warp & 3) == 0) TC_TRACE(s, 4);
        tc_fence_after();
        const uint32_t ta = lane_base + (uint32_t)(q * a_cols);
        if (!(P.debug & 4)) tmem_st32(ta, x);
        mbar_arrive_warp(raw_empty(st), lane);  // staging tile consumed (the TMEM store has read the registers)
        if (split && !(P.debug & 4)) {
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
        if ((warp & 3) == 0) TC_TRACE(s, 5);
        mbar_arrive_warp(a_full(q), lane);
        if ((warp & 3) == 0) TC_TRACE(s, 6);
      }
    }

    // =============================== epilogue ===============================
    // Every lane polls the barrier in its own (inline-asm) loop, so lanes may leave it in different iterations and the
    // warp is NOT guaranteed to be converged afterwards — but everything below is .sync.aligned (tcgen05.ld), which
    // requires the whole warp to execute it together.  Re-converge explicitly.  (Without this the kernel ran correctly
    // almost always and faulted once in a few thousand launches, depending on timing: the intermittent device fault
    // of round 1.)
    mbar_wait(tmem_full, 0);
    __syncwarp();
    tc_fence_after();
    if (threadIdx.x == 0) TC_TRACE_G(2);
    const int quarter = warp & 3;      // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int third = warp >> 2;       // warps w, w+4, w+8 share a quarter and interleave the 16-column groups
    const int row = quarter * 32 + lane;  // TMEM lane == tile row
    const int64_t m = m0 + row;
    const bool row_ok = m < P.M;
    const int64_t mm = row_ok ? m : 0;
    const uint32_t mt = fast_div((uint32_t)mm, P.magic_w, P.shift_w);
    const int w = (int)((uint32_t)mm - mt * (uint32_t)d.grid_w);
    const uint32_t nn_ = fast_div(mt, P.magic_h, P.shift_h);
    const int h = (int)(mt - nn_ * (uint32_t)d.grid_h);
    const int64_t n = (int64_t)nn_;
    const int oh = h * d.out_step + d.out_off_y, ow = w * d.out_step + d.out_off_x;
    const int64_t opix = (n * d.out_h + oh) * (int64_t)d.out_w + ow;
    const bool shuffle = (d.flags & PCODEC_FLAG_PIXEL_SHUFFLE2) != 0;
    const bool has_r2 = d.r2 != nullptr;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    // which accumulators were written (short reductions touch fewer than n_hi_acc hi accumulators)
    uint32_t acc_mask = 0;
    for (int j = 0; j < P.n_hi_acc; ++j)
      if (n_steps > j) acc_mask |= 1u << j;
    if (split && !P.shared_lo) acc_mask |= 1u << P.n_hi_acc;
    if (d.flags & PCODEC_FLAG_SUBPIXEL_NCHW) {
      // Image layer: conv channel (2*py + px) * C + c  ->  out[n][c][2h + py][2w + px] (NCHW, out_h x out_w = 2 x grid).
      // A thread owns pixel (h, w); consecutive lanes are consecutive w, so each (c, py) is one 8-byte store per lane
      // and 256 contiguous bytes per warp.  cout <= 16: a single 16-column chunk, done by warps 0-3.
      const int Cimg = d.out_pixel_stride;  // image channels (3)
      if (third == 0) {
        float acc[16];
        tmem_ld16(lane_addr, acc);
        for (int a = 1; a < n_acc; ++a) {
          if (!((acc_mask >> a) & 1u)) continue;
          float part[16];
          tmem_ld16(lane_addr + (uint32_t)(a * bn), part);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += part[j];
        }
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
              for (int j = 0; j < 16; ++j) {  // static register indexing
                if (j == (2 * py) * Cimg + c) v0 = acc[j];
                if (j == (2 * py + 1) * Cimg + c) v1 = acc[j];
              }
              float *dst = d.out + (n * Cimg + c) * plane + (int64_t)(2 * h + py) * d.out_w + 2 * w;
              *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
            }
        }
      }
    } else if (!shuffle) {
      // Coalesced epilogue: a thread owns one tile ROW in tensor memory, but rows are `out_pixel_stride` floats
      // apart in global memory, so a row-per-thread store touches 32 lines per instruction.  Each warp therefore
      // transposes its 32 rows x 16 columns through a private 2 KB shared-memory tile (the pipeline stages are
      // idle by now; 16-byte chunks XOR-swizzled so both phases are bank-conflict free) and then works with
      // 4 lanes per row / 8 rows per instruction: residual loads and output stores are full 64-byte runs.
      uint8_t *stg = smem_gen + warp * 2048;
      const int rl = lane >> 2, cc = lane & 3;  // transposed phase: row (within a group of 8) and 16-byte chunk
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
      if (threadIdx.x == 0) TC_TRACE_G(9);
      auto load_bias = [&](int c0) {
        const int co = n0 + c0 + 4 * cc;
        return (d.bias && co < d.cout) ? __ldg(reinterpret_cast<const float4 *>(d.bias + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      const bool two = n_acc > 1 && ((acc_mask >> 1) & 1u);
      for (int c0 = third * 16; c0 < bn; c0 += 16 * (TC_PRODUCER_WARPS / 4)) {
        if (n0 + c0 >= d.cout) break;  // padded last N tile
        const int co = n0 + c0 + 4 * cc;
        // What does not depend on the accumulators first: the bias and the first pass's residual values (L2 hits,
        // prefetched at kernel start) are in flight while tensor memory is read, the second pass's during the first.
        float4 bias4;
        if (NG == 2) bias4 = load_bias(c0);
        auto load_res1 = [&](int i) {
          return (d.r1 && ((ok_t >> i) & 1u)) ? __ldg(reinterpret_cast<const float4 *>(d.r1 + opix_t[i] * d.r1_pixel_stride + co))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        const float4 z4_ = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ra0 = NG == 2 ? load_res1(0) : z4_, ra1 = NG == 2 ? load_res1(1) : z4_;  // (NG == 1: no registers to spare)
        __syncwarp();  // the predicated residual loads above may have split the warp; the TMEM loads are .sync.aligned
        float acc[16];
        if (NG != 2) {  // (96-register variants)
          tmem_ld16(lane_addr + (uint32_t)c0, acc);  // warp-collective
          if (two) {
            float part[16];
            tmem_ld16(lane_addr + (uint32_t)(bn + c0), part);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += part[j];
          }
        } else {  // two accumulators per wait
          uint32_t t0[16], t1[16];
          tmem_ld16_issue(lane_addr + (uint32_t)c0, t0);
          if (two) tmem_ld16_issue(lane_addr + (uint32_t)(bn + c0), t1);
          tmem_wait_ld();
          tmem_pin(t0);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(t0[j]);
          if (two) {
            tmem_pin(t1);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(t1[j]);
          }
        }
#pragma unroll 1
        for (int a = 2; a < n_acc; ++a) {
          if (!((acc_mask >> a) & 1u)) continue;
          float part[16];
          tmem_ld16(lane_addr + (uint32_t)(a * bn + c0), part);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += part[j];
        }
        if (threadIdx.x == 0 && c0 == 0) TC_TRACE_G(5);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4 *>(stg + wr_off + (((uint32_t)q ^ wr_x) << 4)) =
              make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        __syncwarp();
        if (threadIdx.x == 0 && c0 == 0) TC_TRACE_G(6);
        if (NG != 2) bias4 = load_bias(c0);
        const float4 rb0 = NG == 2 ? load_res1(2) : z4_, rb1 = NG == 2 ? load_res1(3) : z4_;
#pragma unroll
        for (int i = 0; i < 4; i += 2) {  // two tile rows (i, i + 1) per call
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          F8 v, a1, a2;
          {
            const int r0 = i * 8 + rl, r1_ = (i + 1) * 8 + rl;
            v.a = *reinterpret_cast<const float4 *>(stg + r0 * 64 + (((uint32_t)cc ^ ((uint32_t)(r0 >> 1) & 3u)) << 4));
            v.b = *reinterpret_cast<const float4 *>(stg + r1_ * 64 + (((uint32_t)cc ^ ((uint32_t)(r1_ >> 1) & 3u)) << 4));
          }
          const bool ok0 = (ok_t >> i) & 1u, ok1 = (ok_t >> (i + 1)) & 1u;
          a1.a = NG == 2 ? (i == 0 ? ra0 : rb0) : load_res1(i);
          a1.b = NG == 2 ? (i == 0 ? ra1 : rb1) : load_res1(i + 1);
          a2.a = (d.r2 && ok0) ? __ldg(reinterpret_cast<const float4 *>(d.r2 + opix_t[i] * d.r2_pixel_stride + co)) : z4;
          a2.b = (d.r2 && ok1) ? __ldg(reinterpret_cast<const float4 *>(d.r2 + opix_t[i + 1] * d.r2_pixel_stride + co)) : z4;
          v.a = make_float4(v.a.x + bias4.x, v.a.y + bias4.y, v.a.z + bias4.z, v.a.w + bias4.w);
          v.b = make_float4(v.b.x + bias4.x, v.b.y + bias4.y, v.b.z + bias4.z, v.b.w + bias4.w);
          if (threadIdx.x == 0 && c0 == 0 && i == 0) TC_TRACE_G(10);
          const F8 o = tc_epilogue8(d.epilogue, v, a1, a2, has_r2);
          if (threadIdx.x == 0 && c0 == 0 && i == 0 && o.a.x != 123.f) TC_TRACE_G(11);
          if (ok0) *reinterpret_cast<float4 *>(d.out + opix_t[i] * d.out_pixel_stride + co) = o.a;
          if (ok1) *reinterpret_cast<float4 *>(d.out + opix_t[i + 1] * d.out_pixel_stride + co) = o.b;
          if (threadIdx.x == 0 && c0 == 0 && i == 0) TC_TRACE_G(7);
        }
        if (threadIdx.x == 0 && c0 == 0) TC_TRACE_G(8);
        __syncwarp();
      }
    } else {
      for (int c0 = third * 16; c0 < bn; c0 += 16 * (TC_PRODUCER_WARPS / 4)) {
        if (n0 + c0 >= d.cout) break;  // padded last N tile
        __syncwarp();  // lanes that skipped the stores of the previous round (`continue`) rejoin before the collective load
        float acc[16];
        tmem_ld16(lane_addr + (uint32_t)c0, acc);  // warp-collective: executed by all lanes, stores are predicated
        for (int a = 1; a < n_acc; ++a) {
          if (!((acc_mask >> a) & 1u)) continue;
          float part[16];
          tmem_ld16(lane_addr + (uint32_t)(a * bn + c0), part);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += part[j];
        }
        if (!row_ok) continue;
        const int co0 = n0 + c0;
#pragma unroll 1
        for (int j4 = 0; j4 < 4; ++j4) {  // 4 consecutive conv channels = the 2x2 sub-pixels of one output channel
          const int co = co0 + 4 * j4;
          const float4 b4 = d.bias ? __ldg(reinterpret_cast<const float4 *>(d.bias + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float a4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) a4[e] = acc[0];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)  // static register indexing
            if ((jj >> 2) == j4) a4[jj & 3] = acc[jj];
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 o = tc_epilogue4(d.epilogue, make_float4(a4[0] + b4.x, a4[1] + b4.y, a4[2] + b4.z, a4[3] + b4.w), z, z, false);
          const int c = co >> 2;
          const int64_t sp0 = (n * d.out_h + 2 * oh) * (int64_t)d.out_w + 2 * ow;
          float *o0 = d.out + sp0 * d.out_pixel_stride + c;
          o0[0] = o.x;
          o0[d.out_pixel_stride] = o.y;
          o0[(int64_t)d.out_w * d.out_pixel_stride] = o.z;
          o0[((int64_t)d.out_w + 1) * d.out_pixel_stride] = o.w;
        }
      }
    }
    tc_fence_before();
    if (threadIdx.x == 0) TC_TRACE_G(3);
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
        if (P.debug & 2) {
          mbar_arrive(full_b(st));
        } else {
          mbar_expect_tx(full_b(st), (uint32_t)((split ? 2 : 1) * b_bytes));
          tma_load_2d(b_hi(st), &map_hi, full_b(st), k, n0);
          if (split) tma_load_2d(b_lo(st), &map_lo, full_b(st), k, n0);
        }
        const int sc = d.seg[seg].channels;
        if (++kc == (sc + TC_BK - 1) / TC_BK) {
          kc = 0;
          if (++tap == d.n_taps) { tap = 0; seg_cbase += sc; ++seg; }
        }
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    // The whole warp runs the loop (uniform control flow) and one lane chosen with elect.sync issues: with
    // `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop (~55 clk per instruction, which made
    // the issuing thread the bottleneck); elected in a converged warp the 12 MMAs of a slab issue straight-line at
    // the tensor pipe's own rate (N/2 clk each — tools/ubench/mma_rate.cu).
    {
      // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t acc_lo = P.shared_lo ? tmem_acc : tmem_acc + (uint32_t)(P.n_hi_acc * bn);
      int st = 0, hi_idx = 0, q = 0;
      uint32_t qph = 0;
      // tcgen05.commit returns only when the queued MMAs have (nearly) retired, so every cycle between a commit and
      // the next slab's first MMA is tensor-pipe idle time: the whole loop therefore runs in ONE elected thread (no
      // per-slab __syncwarp / re-election; a poll of an already-completed mbarrier by the issuing thread itself costs
      // ~15 clk against ~135 clk for lane-0-polls-then-syncwarp).
      if (elect_one()) {
        for (int s = 0; s < n_steps; ++s) {
          TC_TRACE(s, 7);
          mbar_wait(a_full(q), qph);  // A operand in TMEM and (checked by the converters) B in shared memory
          TC_TRACE(s, 8);
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
              umma_tf32_ts(acc_lo, ta_lo + ak, db_hi + adv, idesc, (P.shared_lo || s > 0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(acc_lo, ta_hi + ak, db_lo + adv, idesc, 1u);
            }
          }
          TC_TRACE(s, 11);
          umma_commit(empty_b(st));  // B slot + A operand buffer reusable once these MMAs retire
          TC_TRACE(s, 10);
          if (++st == stages) st = 0;
          if (++hi_idx == P.n_hi_acc) hi_idx = 0;
          if (++q == a_ring) { q = 0; qph ^= 1u; }
        }
        umma_commit(tmem_full);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, tmem_cols);
  }
  if (threadIdx.x == 0) TC_TRACE_G(4);
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

// N tile: the fewest tiles (least A-operand re-reading, fewest CTAs) such that the accumulators fit tensor memory
// next to a 2-deep A operand ring AND no hi accumulator sees more than ~320 MMAs: the tensor core truncates once per
// MMA per accumulator, so spreading a long reduction over several accumulators keeps the result at fp32-class
// accuracy (measured: rms 3e-5 -> 3e-6 at K = 4800 going from 1 to 4 accumulators).  Tiles are multiples of 16 and
// need not divide cout: the last tile may be padded (TMA zero-fills out-of-range weight rows, the epilogue skips
// columns >= cout).
int pick_bn(int cout, int k_total, int cap = 256, int tmem_cols = 512) {
  if (cout % 16 != 0) return 0;
  if (const char *e = pcodec_knob("PCODEC_TC_BNCAP")) cap = atoi(e);  // experiment knob
  const int n_steps = (k_total + TC_BK - 1) / TC_BK;
  const int need_hi = std::max(1, (4 * n_steps + 319) / 320);
  for (int tiles = 1; tiles <= 64; ++tiles) {
    int bn = (cout + tiles - 1) / tiles;
    bn = (bn + 15) & ~15;
    if (bn > cap) continue;
    const int n_acc = (tmem_cols - 128) / bn;  // next to two 64-column A operand buffers
    if (n_acc - 1 >= std::min(need_hi, 4) || bn == 16) return bn;
  }
  return 0;
}

}  // namespace

extern "C" int pcodec_conv_tc_prepare(const float *w_tap_major, int n_taps, int cin_total, int cout, void **handle_out,
                                      void *stream) {
  if (!w_tap_major || !handle_out || n_taps < 1 || cin_total < 4 || cout < 1) return PCODEC_ERR_BAD_ARG;
  *handle_out = nullptr;
  const int n_slabs = (n_taps * cin_total + TC_BK - 1) / TC_BK;
  // Default: only 1-tap reductions of <= 24 slabs whose whole output fits ONE tile of <= 128 columns (the ResidualUnit
  // 1x1 192->96: -23 %).  With so few MMAs (<= 288) the lo products can share the hi accumulator, so the tile needs
  // bn + 128 <= 256 TMEM columns and ~82 KB of shared memory, and two CTAs overlap each other's prologue / epilogue.
  // For wider outputs the variant needs more N tiles (more A re-reads) and measured equal or slower; for the narrow 3x3
  // slice-stack layers two resident CTAs did not raise the per-SM throughput.  PCODEC_TC_SMALL: 0 = never, 1 = also
  // every cout <= 64 layer, 2 = also every 1-tap short reduction (separate lo accumulator, bn <= 64), 3 = as 2 with the
  // shared accumulator (bn <= 128).
  int mode = -1;
  if (const char *e = pcodec_knob("PCODEC_TC_SMALL")) mode = atoi(e);
  const bool short_1tap = n_taps == 1 && n_slabs <= 24;
  bool small = mode < 0 ? (short_1tap && cout <= 128) : ((mode >= 1 && cout <= 64) || (mode >= 2 && short_1tap));
  const bool shared = small && short_1tap && (mode < 0 || mode >= 3);
  // mode 4: tiles of up to 192 columns in the two-CTA variant (192 accumulator columns + ONE 64-column A buffer = 256,
  // one pipeline stage): the main loop is serialised, but it is short and hides behind the other CTA's epilogue.
  const bool wide = shared && mode == 4;
  const int bn = small ? pick_bn(cout, n_taps * cin_total, wide ? 192 : (shared ? 128 : 64), wide ? 512 : (shared ? 256 + 128 : 256))
                       : pick_bn(cout, n_taps * cin_total);
  if (bn == 0 || (cin_total % 4) != 0) return PCODEC_ERR_UNSUPPORTED;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PCODEC_ERR_UNSUPPORTED;
  TcWeights *h = new TcWeights();
  h->cout = cout;
  h->k_total = n_taps * cin_total;
  h->bn = bn;
  h->n_tiles = (cout + bn - 1) / bn;
  h->small = small ? (shared ? 2 : 1) : 0;
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
  if (e != cudaSuccess) {
    cudaFree(h->dev_hi);
    cudaFree(h->dev_lo);
    delete h;
    return -(int)e;
  }
  h->w16 = pcodec_tc16_weights_create(w_tap_major, n_taps, cin_total, cout, stream);
  *handle_out = h;
  return PCODEC_OK;
}

extern "C" int pcodec_debug_tc_trace(long long *out, int n) {
  const int total = TC_TRACE_SLABS * TC_TRACE_EVENTS + 32;
  if (!out || n < total) return total;
  return cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * total) == cudaSuccess ? total : -1;
}

extern "C" void pcodec_conv_tc_release(void *handle) {
  if (!handle) return;
  TcWeights *h = static_cast<TcWeights *>(handle);
  pcodec_tc16_weights_destroy(h->w16);
  cudaFree(h->dev_hi);
  cudaFree(h->dev_lo);
  delete h;
}

const void *pcodec_tc_w16(const void *tc_weights) {
  return tc_weights ? static_cast<const TcWeights *>(tc_weights)->w16 : nullptr;
}

bool pcodec_conv_taps_tc_supported(const pcodec_conv_desc *d) {
  if (!d->tc_weights) return false;
  const TcWeights *h = static_cast<const TcWeights *>(d->tc_weights);
  if (h->cout != d->cout || h->k_total != d->n_taps * d->cin_total) return false;
  if (d->tc_split != 1 && d->tc_split != 3) return false;
  if (d->flags & PCODEC_FLAG_SUBPIXEL_NCHW)  // out_pixel_stride = image channels; float2 stores
    return d->cout == 16 && d->out_pixel_stride >= 1 && d->out_pixel_stride <= 4 && 4 * d->out_pixel_stride <= 16 &&
           (reinterpret_cast<uintptr_t>(d->out) & 7) == 0 && (d->out_w & 1) == 0;
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
  if (P.M >= (1ll << 31) - TC_BM) return PCODEC_ERR_UNSUPPORTED;  // 32-bit pixel arithmetic in the kernel
  auto magic_for = [](uint32_t dv, uint32_t &magic, int &shift) {
    shift = 0;
    while ((1u << shift) < dv) ++shift;
    magic = (uint32_t)((((uint64_t)1 << (31 + shift)) + dv - 1) / dv);
  };
  magic_for((uint32_t)desc->grid_w, P.magic_w, P.shift_w);
  magic_for((uint32_t)desc->grid_h, P.magic_h, P.shift_h);
  P.bn = h->bn;
  P.split = desc->tc_split;
  int n_steps = 0;
  for (int s = 0; s < desc->n_segments; ++s) n_steps += desc->n_taps * ((desc->seg[s].channels + TC_BK - 1) / TC_BK);
  P.n_steps = n_steps;
  const bool split3 = P.split == 3;
  const int stage_bytes = TC_A_BYTES + (split3 ? 2 : 1) * h->bn * 128;  // raw A staging tile | B_hi | (B_lo)
  auto need = [&](int st) { return std::max(st * stage_bytes, 12 * 2048) + 1024 + 8 * (4 * st + 6) + 64; };
  const int smem_limit = h->small ? 110 * 1024 : TC_SMEM_LIMIT;
  const int tmem_limit = h->small ? 256 : 512;
  int stages = 1;
  while (need(stages + 1) <= smem_limit && stages < 8) ++stages;
  if (const char *e = pcodec_knob("PCODEC_TC_STAGES")) stages = std::min(stages, atoi(e));  // experiment knob
  if (stages > n_steps) stages = n_steps;
  if (stages < 1 || need(stages) > smem_limit) return PCODEC_ERR_UNSUPPORTED;
  P.stages = stages;
  P.debug = 0;
  if (const char *e = pcodec_knob("PCODEC_TC_DEBUG")) P.debug = atoi(e);  // experiment knob (wrong results!)
  {
    // TMEM columns: (n_hi hi accumulators + 1 lo) of bn columns + a_ring A operand buffers (hi 32 [+ lo 32] columns).
    // Prefer a 3-deep A ring; spend what is left on hi accumulators (up to 4).
    const int a_cols = split3 ? 64 : 32;
    const int need_hi = std::min(4, std::max(1, (4 * n_steps + 319) / 320));  // same rule as pick_bn
    int ring = h->small ? 2 : 3;
    int n_acc = (tmem_limit - ring * a_cols) / h->bn;
    if (n_acc - (split3 ? 1 : 0) < need_hi) {  // wide tile: a 2-deep ring rather than too few hi accumulators
      ring = 2;
      n_acc = (tmem_limit - ring * a_cols) / h->bn;
    }
    if (const char *e = pcodec_knob("PCODEC_TC_RING")) {  // experiment knob
      ring = std::max(2, std::min(4, atoi(e)));
      n_acc = (tmem_limit - ring * a_cols) / h->bn;
    }
    P.shared_lo = 0;
    if (h->small == 2) {  // single shared accumulator
      P.shared_lo = 1;
      n_acc = 1 + (split3 ? 1 : 0);
      if (h->bn + ring * a_cols > tmem_limit) ring = 1;
    }
    int n_hi = n_acc - (split3 ? 1 : 0);
    if (n_hi > 4) n_hi = 4;
    if (n_hi < 1) return PCODEC_ERR_UNSUPPORTED;
    if (ring > stages) ring = stages;  // the "consumed" barrier of slab s - ring must not have been recycled
    if (ring < 1) ring = 1;
    P.a_ring = ring;
    P.n_hi_acc = n_hi;
  }
  const int smem = need(stages);
  if (getenv("PCODEC_TC_VERBOSE"))
    fprintf(stderr, "[conv_tc] M=%lld bn=%d n_tiles=%d n_steps=%d stages=%d n_hi=%d ring=%d split=%d smem=%d small=%d\n",
            (long long)P.M, h->bn, h->n_tiles, n_steps, stages, P.n_hi_acc, P.a_ring, P.split, smem, h->small);
  {  // opt-in to > 48 KB of dynamic shared memory, once per device (the attribute is per device)
    static std::atomic<uint64_t> attr_mask{0};
    uint64_t bit;
    if (pcodec_device_needs(attr_mask, &bit)) {
      PCODEC_CHECK_CUDA(cudaFuncSetAttribute(conv_taps_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
      PCODEC_CHECK_CUDA(cudaFuncSetAttribute(conv_taps_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      attr_mask.fetch_or(bit, std::memory_order_release);
    }
  }
  dim3 grid((unsigned)ceil_div64(P.M, TC_BM), (unsigned)h->n_tiles);
  // PCODEC_TC_NG=3: a third converter group (18 warps).  EXPERIMENT, not yet run on hardware: with every load and TMEM
  // store switched off the slab period is still ~760 clk, i.e. the two groups' wait -> LDS -> store -> arrive cycle is
  // the floor of the long reductions (lesson 6).
  static const bool three_groups = [] { const char *e = pcodec_knob("PCODEC_TC_NG"); return e && atoi(e) == 3; }();
  if (h->small)
    conv_taps_tc_kernel<1><<<grid, 32 * 10, smem, as_stream(stream)>>>(P, h->map_hi, h->map_lo);
  else if (three_groups && smem >= 16 * 2048 + 2048) {
    static const cudaError_t attr3 =
        cudaFuncSetAttribute(conv_taps_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (attr3 != cudaSuccess) return -(int)attr3;
    conv_taps_tc_kernel<3><<<grid, 32 * 18, smem, as_stream(stream)>>>(P, h->map_hi, h->map_lo);
  } else
    conv_taps_tc_kernel<2><<<grid, 32 * 14, smem, as_stream(stream)>>>(P, h->map_hi, h->map_lo);
  PCODEC_RETURN_LAUNCH();
}
