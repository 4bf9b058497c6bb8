// tcgen05 / TMEM / TMA / mbarrier PTX wrappers shared by the tensor-core convolution kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "common.cuh"

namespace pcodec_tc {

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
// exact m / d for m < 2^31 with magic = ceil(2^(31+shift) / d), shift = ceil(log2 d)  (Granlund-Montgomery, N = 31):
// the tile's pixel coordinates cost 2 multiplies per row instead of 2 integer divisions (~40 instructions each), which
// was ~2 us of every tile's prologue
__device__ __forceinline__ uint32_t fast_div(uint32_t m, uint32_t magic, int shift) {
  return (uint32_t)(((uint64_t)m * magic) >> (31 + shift));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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

// One lane polls the barrier, the warp converges behind it: an mbarrier op per THREAD (128 try_waits + 128 arrives
// per barrier per K slab) serialises in the shared-memory unit and was the whole per-slab cost of the v5 pipeline
// (measured: ~1200 clk per slab with every load, TMEM store and 2 of 3 MMAs removed).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// one lane of a CONVERGED warp (PTX elect.sync): lets ptxas emit the single-thread tcgen05 instructions straight-line
// instead of wrapping each in an ELECT / BRA.U.ANY loop (what `if (lane == 0)` compiles to)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, M=128, N from idesc, K=8 (TS form): A = 128 lanes x 8 columns of fp32 at tmem_a
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

// Asynchronous form: the registers are valid after tmem_wait_ld(); tmem_pin() keeps the compiler from moving their
// uses in front of that wait.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_pin(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
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


// 4-D tiled TMA load (activation planes [N][H][W][C]: coordinates c, w, h, n; out-of-range elements are zero-filled, which
// IS the convolution's padding)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 operands, fp32 accumulate), M=128, N from idesc, K=16 (SS form)
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Four channels at once behind ONE switch: keeps a single copy of every transcendental in the kernel image (the code is
// fetched cold by every CTA; 16 inlined copies of the scalar switch made the kernel 160 KB of SASS) and lets the four
// independent chains of a case interleave (the epilogue warps are latency bound, not throughput bound).
static __device__ __noinline__ float4 tc_epilogue4(int epi, float4 v, float4 r1, float4 r2, bool has_r2) {
#define PC_EACH(expr)                                                                  \
  do {                                                                                 \
    { const float a = v.x, p = r1.x, q = r2.x; (void)p; (void)q; o.x = (expr); }        \
    { const float a = v.y, p = r1.y, q = r2.y; (void)p; (void)q; o.y = (expr); }        \
    { const float a = v.z, p = r1.z, q = r2.z; (void)p; (void)q; o.z = (expr); }        \
    { const float a = v.w, p = r1.w, q = r2.w; (void)p; (void)q; o.w = (expr); }        \
  } while (0)
  float4 o;
  switch (epi) {
    case PCODEC_EPI_GELU:
      gelu_erf2(v.x, v.y, o.x, o.y);
      gelu_erf2(v.z, v.w, o.z, o.w);
      break;
    case PCODEC_EPI_ADD: PC_EACH(a + p); break;
    case PCODEC_EPI_ADD_GELU:
      gelu_erf2(v.x + r1.x, v.y + r1.y, o.x, o.y);
      gelu_erf2(v.z + r1.z, v.w + r1.w, o.z, o.w);
      break;
    case PCODEC_EPI_GATE: PC_EACH(q * sigmoid_f(a) + p); break;
    case PCODEC_EPI_GDN: PC_EACH(p * rsqrtf(a)); break;
    case PCODEC_EPI_IGDN: PC_EACH(p * sqrtf(a)); break;
    case PCODEC_EPI_LRP:
      if (has_r2) PC_EACH(__fadd_rn(__fadd_rn(p, __fmul_rn(0.5f, tanhf(a))), q));
      else PC_EACH(__fadd_rn(p, __fmul_rn(0.5f, tanhf(a))));
      break;
    case PCODEC_EPI_CLAMP01: PC_EACH(fminf(fmaxf(a, 0.f), 1.f)); break;
    case PCODEC_EPI_LEAKY: PC_EACH(a > 0.f ? a : __fmul_rn(0.01f, a)); break;
    case PCODEC_EPI_LEAKY_ADD: PC_EACH((a > 0.f ? a : __fmul_rn(0.01f, a)) + p); break;
    default: o = v; break;
  }
#undef PC_EACH
  return o;
}

// Eight values (two float4 of two tile rows) per call: twice the independent chains per warp for the latency-bound
// epilogue warps.
struct F8 { float4 a, b; };
static __device__ __noinline__ F8 tc_epilogue8(int epi, F8 v, F8 r1, F8 r2, bool has_r2) {
#define PC_EACH8(expr)                                                                          \
  do {                                                                                         \
    { const float a = v.a.x, p = r1.a.x, q = r2.a.x; (void)p; (void)q; o.a.x = (expr); }        \
    { const float a = v.b.x, p = r1.b.x, q = r2.b.x; (void)p; (void)q; o.b.x = (expr); }        \
    { const float a = v.a.y, p = r1.a.y, q = r2.a.y; (void)p; (void)q; o.a.y = (expr); }        \
    { const float a = v.b.y, p = r1.b.y, q = r2.b.y; (void)p; (void)q; o.b.y = (expr); }        \
    { const float a = v.a.z, p = r1.a.z, q = r2.a.z; (void)p; (void)q; o.a.z = (expr); }        \
    { const float a = v.b.z, p = r1.b.z, q = r2.b.z; (void)p; (void)q; o.b.z = (expr); }        \
    { const float a = v.a.w, p = r1.a.w, q = r2.a.w; (void)p; (void)q; o.a.w = (expr); }        \
    { const float a = v.b.w, p = r1.b.w, q = r2.b.w; (void)p; (void)q; o.b.w = (expr); }        \
  } while (0)
  F8 o;
  switch (epi) {
    case PCODEC_EPI_GELU:
      gelu_erf2(v.a.x, v.a.y, o.a.x, o.a.y);
      gelu_erf2(v.a.z, v.a.w, o.a.z, o.a.w);
      gelu_erf2(v.b.x, v.b.y, o.b.x, o.b.y);
      gelu_erf2(v.b.z, v.b.w, o.b.z, o.b.w);
      break;
    case PCODEC_EPI_ADD: PC_EACH8(a + p); break;
    case PCODEC_EPI_ADD_GELU:
      gelu_erf2(v.a.x + r1.a.x, v.a.y + r1.a.y, o.a.x, o.a.y);
      gelu_erf2(v.a.z + r1.a.z, v.a.w + r1.a.w, o.a.z, o.a.w);
      gelu_erf2(v.b.x + r1.b.x, v.b.y + r1.b.y, o.b.x, o.b.y);
      gelu_erf2(v.b.z + r1.b.z, v.b.w + r1.b.w, o.b.z, o.b.w);
      break;
    case PCODEC_EPI_GATE: PC_EACH8(q * sigmoid_f(a) + p); break;
    case PCODEC_EPI_GDN: PC_EACH8(p * rsqrtf(a)); break;
    case PCODEC_EPI_IGDN: PC_EACH8(p * sqrtf(a)); break;
    case PCODEC_EPI_LRP:
      if (has_r2) PC_EACH8(__fadd_rn(__fadd_rn(p, __fmul_rn(0.5f, tanhf(a))), q));
      else PC_EACH8(__fadd_rn(p, __fmul_rn(0.5f, tanhf(a))));
      break;
    case PCODEC_EPI_CLAMP01: PC_EACH8(fminf(fmaxf(a, 0.f), 1.f)); break;
    case PCODEC_EPI_LEAKY: PC_EACH8(a > 0.f ? a : __fmul_rn(0.01f, a)); break;
    case PCODEC_EPI_LEAKY_ADD: PC_EACH8((a > 0.f ? a : __fmul_rn(0.01f, a)) + p); break;
    default: o = v; break;
  }
#undef PC_EACH8
  return o;
}

// D[tmem] (+)= A[tmem] * B[smem], kind::f16, M=128, K=16 (TS form): A = 128 lanes x 8 columns (two fp16 per column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared memory -> tensor memory, 128 rows x 256 bits (one K = 16 step of an fp16 A operand): row r -> lane r, 8 columns.
// Executes in issue order with the tcgen05.mma of the same thread.
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t tmem_dst, uint64_t smem_desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}

}  // namespace pcodec_tc
