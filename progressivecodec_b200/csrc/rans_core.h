// rANS state arithmetic shared by the device kernels and by host-side unit tests (tests compile this
// header with g++ through pcodec_rans_selftest_* in rans.cu's host section).
//
// Format (bit-exact with the reference coder, /root/reference/src/third_party/ryg_rans/rans64.h and
// compress/cpp_exts/rans/rans_interface.cpp): 64-bit state, lower bound 2^31, 32-bit word
// renormalisation, 16-bit probability precision, 4-bit raw "bypass" tokens.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PC_HD __host__ __device__ __forceinline__
#else
#define PC_HD inline
#endif

namespace pcodec {

constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr uint32_t kBypassMax = 15;
constexpr uint64_t kRansLower = 1ull << 31;

// x / freq and x % freq for x < 2^63, 0 < freq < 2^16, using three 32-bit divisions (a 64-bit
// division is a long software routine on the GPU and sits on the coder's serial dependency chain).
PC_HD void divmod_u63_u16(uint64_t x, uint32_t freq, uint64_t &q, uint32_t &r) {
  uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
  uint32_t q2 = hi / freq;
  uint32_t rem = hi - q2 * freq;                 // < 2^16
  uint32_t t = (rem << 16) | (lo >> 16);         // < freq * 2^16 <= 2^32
  uint32_t q1 = t / freq;                        // < 2^16
  rem = t - q1 * freq;
  t = (rem << 16) | (lo & 0xFFFFu);
  uint32_t q0 = t / freq;                        // < 2^16
  r = t - q0 * freq;
  q = ((uint64_t)q2 << 32) | ((uint64_t)q1 << 16) | (uint64_t)q0;
}

// Encoder: push one modelled symbol. Returns true when a 32-bit word must be emitted (word is then
// valid); rans64.h:77-93 (Rans64EncPut).
PC_HD bool enc_put(uint64_t &x, uint32_t start, uint32_t freq, uint32_t &word) {
  bool emit = false;
  uint64_t x_max = (uint64_t)freq << (31 - kPrecision + 32);
  if (x >= x_max) {
    word = (uint32_t)x;
    x >>= 32;
    emit = true;
  }
  uint64_t q;
  uint32_t r;
  divmod_u63_u16(x, freq, q, r);
  x = (q << kPrecision) + r + start;
  return emit;
}

// ---- division-free variant (Alverson reciprocal, cf. rans64.h Rans64EncSymbolInit/Rans64EncPutSymbol) ----
// For 2 <= freq < 2^16: shift = ceil(log2 freq), rcp = ceil(2^(shift+63)/freq); then for every x < 2^63
//   x / freq == mulhi64(x, rcp) >> (shift - 1), and C(s,x) = x + start + (x/freq) * (65536 - freq).
PC_HD uint32_t rcp_shift_for(uint32_t freq) {  // ceil(log2 freq) - 1, freq >= 2
  uint32_t shift = 0;
  while (freq > (1u << shift)) ++shift;
  return shift - 1;
}
PC_HD uint64_t rcp_for(uint32_t freq) {
  if (freq < 2) return ~0ull;
  const uint32_t shift = rcp_shift_for(freq) + 1;
  const unsigned __int128 num = ((unsigned __int128)1 << (shift + 63)) + freq - 1;
  return (uint64_t)(num / freq);
}
PC_HD uint64_t mulhi_u64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
PC_HD bool enc_put_rcp(uint64_t &x, uint32_t start, uint32_t freq, uint64_t rcp, uint32_t rshift, uint32_t &word) {
  const uint64_t x_max = (uint64_t)freq << (31 - kPrecision + 32);
  const bool emit = x >= x_max;
  word = (uint32_t)x;
  if (emit) x >>= 32;
  const uint64_t q = freq >= 2 ? (mulhi_u64(x, rcp) >> rshift) : x;
  x = x + start + q * (uint64_t)(65536u - freq);
  return emit;
}

// Encoder: push one raw 4-bit value; rans_interface.cpp:60-78 (Rans64EncPutBits, nbits = 4).
PC_HD bool enc_put_bits4(uint64_t &x, uint32_t val, uint32_t &word) {
  bool emit = false;
  const uint64_t x_max = 1ull << (31 - 16 + 32 + (16 - kBypassBits));
  if (x >= x_max) {
    word = (uint32_t)x;
    x >>= 32;
    emit = true;
  }
  x = (x << kBypassBits) | val;
  return emit;
}

// Decoder: advance past a symbol (start,freq); returns true when a refill word must be consumed
// (caller then does x = (x << 32) | word); rans64.h:126-142 (Rans64DecAdvance).
PC_HD bool dec_advance(uint64_t &x, uint32_t start, uint32_t freq) {
  x = (uint64_t)freq * (x >> kPrecision) + (x & 0xFFFFu) - start;
  return x < kRansLower;
}

// Decoder: pop 4 raw bits; rans_interface.cpp:80-96. Returns true when a refill is needed.
PC_HD bool dec_get_bits4(uint64_t &x, uint32_t &val) {
  val = (uint32_t)(x & kBypassMax);
  x >>= kBypassBits;
  return x < kRansLower;
}

// Symbol -> (table slot, raw escape payload); rans_interface.cpp:114-128.
PC_HD void classify(int32_t symbol, int32_t offset, int32_t max_value, int32_t &slot, uint32_t &raw, bool &esc) {
  int32_t v = symbol - offset;
  raw = 0;
  if (v < 0) {
    raw = (uint32_t)(-2 * v - 1);
    v = max_value;
  } else if (v >= max_value) {
    raw = (uint32_t)(2 * (v - max_value));
    v = max_value;
  }
  slot = v;
  esc = (v == max_value);
}

PC_HD int nibble_count(uint32_t raw) {
  int nb = 0;
  while (nb < 8 && (raw >> (nb * kBypassBits)) != 0) ++nb;
  return nb;
}

}  // namespace pcodec
