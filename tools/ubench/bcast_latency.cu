// Latency of the two ways a warp can broadcast "the one valid lane's 64-bit value" (the rANS decoder's per-symbol step):
//   A: two redux.sync.or over values masked by validity      B: ballot + ffs + two shfl.idx from the winning lane
// Dependent chains (the next iteration's validity depends on the broadcast value), one warp, clock64 per iteration.
#include <cstdio>
#include <cstdint>
__global__ void k(long long *out, int iters) {
  const int lane = threadIdx.x & 31;
  uint32_t lo = 12345u, hi = 777u;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const bool valid = ((lo >> 3) & 31u) == (uint32_t)lane;
    const uint32_t cl = lo * 2654435761u + (uint32_t)lane, ch = hi + (cl >> 7);
    lo = __reduce_or_sync(0xFFFFFFFFu, valid ? cl : 0u);
    hi = __reduce_or_sync(0xFFFFFFFFu, valid ? ch : 0u);
  }
  long long t1 = clock64();
  uint32_t lo2 = 12345u, hi2 = 777u;
  for (int i = 0; i < iters; ++i) {
    const bool valid = ((lo2 >> 3) & 31u) == (uint32_t)lane;
    const uint32_t cl = lo2 * 2654435761u + (uint32_t)lane, ch = hi2 + (cl >> 7);
    const int src = __ffs(__ballot_sync(0xFFFFFFFFu, valid)) - 1;
    lo2 = __shfl_sync(0xFFFFFFFFu, cl, src);
    hi2 = __shfl_sync(0xFFFFFFFFu, ch, src);
  }
  long long t2 = clock64();
  uint32_t lo3 = 12345u, hi3 = 777u;  // C: plain ALU chain of the same arithmetic (no broadcast): the floor
  for (int i = 0; i < iters; ++i) {
    const uint32_t cl = lo3 * 2654435761u + (uint32_t)lane, ch = hi3 + (cl >> 7);
    lo3 = cl; hi3 = ch;
  }
  long long t3 = clock64();
  if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = lo ^ hi ^ lo2 ^ hi2 ^ lo3 ^ hi3; }
}
int main() {
  long long *d, h[4];
  cudaMalloc(&d, 32);
  const int iters = 100000;
  k<<<1, 32>>>(d, iters);
  k<<<1, 32>>>(d, iters);
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("clk per iteration: 2x redux.or %.1f, ballot+ffs+2x shfl %.1f, ALU only %.1f (check %lld)\n", (double)h[0] / iters,
         (double)h[1] / iters, (double)h[2] / iters, h[3]);
  return 0;
}
