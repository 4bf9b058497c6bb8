// Micro-benchmark: issue rate / execution time of tcgen05.mma kind::tf32 (M=128, K=8) for several N, A from TMEM
// (TS) or from shared memory (SS), B from shared memory (SWIZZLE_128B K-major).  Operands are garbage; timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: TS tf32, all MMAs into ONE accumulator; 1: TS, alternate 2 accumulators; 2: SS tf32 one accumulator;
// 3: SS bf16 (K=16) one accumulator; 4: TS tf32, 3-product pattern (acc_hi, acc_lo, acc_lo)
template <int mode>
__global__ void __launch_bounds__(128, 1) k(int N, int n_mma, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar, bar2, bar3;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t da = desc_sw128(base), db = desc_sw128(base + 16384);
    const uint32_t a_tmem = tm + 448;  // A operand columns
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one()) {
      t0 = clock64();
      const uint32_t tmN = tm + N;
      for (int i = 0; mode < 5 && i < n_mma; i += 12) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const uint64_t adv = (uint64_t)((j & 3) * 2);
          if (mode == 0) mma_ts(tm, a_tmem + (j & 3) * 8, db + adv, idesc_tf32, 1);
          else if (mode == 1) mma_ts((j & 1) ? tmN : tm, a_tmem + (j & 3) * 8, db + adv, idesc_tf32, 1);
          else if (mode == 2) mma_ss(tm, da + adv, db + adv, idesc_tf32, 1);
          else if (mode == 3) mma_ss_f16(tm, da + adv, db + adv, idesc_bf16, 1);
          else mma_ts((j % 3) ? tmN : tm, a_tmem + (j & 3) * 8 + ((j % 3) == 1 ? 32 : 0), db + adv + ((j % 3) == 2 ? 1024 : 0), idesc_tf32, 1);
        }
      }
      if (mode >= 5) {
        // kernel-like: per slab [poll a completed barrier] + 12 MMAs (3-product pattern) + commit
        for (int i = 0; i < n_mma; i += 12) {
          if (mode >= 6) {
            asm volatile("{\n\t.reg .pred P1;\n\tWL2:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 1, 0x989680;\n\t@P1 bra WD2;\n\tbra WL2;\n\tWD2:\n\t}" ::"r"(smem_u32(&bar2)) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const uint64_t adv = (uint64_t)((j & 3) * 2);
            mma_ts((j % 3) ? tmN : tm, a_tmem + (j & 3) * 8 + ((j % 3) == 1 ? 32 : 0), db + adv + ((j % 3) == 2 ? 1024 : 0), idesc_tf32, 1);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar3)) : "memory");
          if (mode >= 7)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar3)) : "memory");
        }
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0, 0x989680;\n\t@P1 bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
      t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char *names[] = {"TS tf32 1acc", "TS tf32 2acc", "SS tf32 1acc", "SS bf16 1acc", "TS tf32 3prod", "3prod+commit", "poll+3p+commit", "poll+3p+2commit"};
  for (int mode = 4; mode < 8; ++mode)
    for (int N : {32, 64, 80, 96, 112, 128, 192, 256}) {
      if ((mode == 1 || mode >= 4) && 2 * N > 448) continue;
      if (N > 448) continue;
      const int n = 1200;
      switch (mode) {
        case 0: k<0><<<1, 128, 100 * 1024>>>(N, n, d); break;
        case 1: k<1><<<1, 128, 100 * 1024>>>(N, n, d); break;
        case 2: k<2><<<1, 128, 100 * 1024>>>(N, n, d); break;
        case 3: k<3><<<1, 128, 100 * 1024>>>(N, n, d); break;
        case 4: k<4><<<1, 128, 100 * 1024>>>(N, n, d); break;
        case 5: k<5><<<1, 128, 100 * 1024>>>(N, n, d); break;
        case 6: k<6><<<1, 128, 100 * 1024>>>(N, n, d); break;
        default: k<7><<<1, 128, 100 * 1024>>>(N, n, d); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-14s N=%3d: issue %6.1f clk/MMA, issue+drain %6.1f clk/MMA (ideal %5.1f)\n", names[mode], N, (double)h[0] / n,
             (double)h[1] / n, mode == 3 ? N / 2.0 : N / 2.0);
    }
  return 0;
}
