/*
 * CPU ORACLE (test infrastructure; never linked into the product library).
 *
 * Plain-C restatement of the reference entropy coder:
 *   - token expansion incl. 4-bit bypass escapes:  rans_interface.cpp:99-164
 *   - reverse-order rANS flush, 64-bit state, 32-bit renorm, 16-bit precision:
 *         rans_interface.cpp:166-191 + ryg_rans/rans64.h:77-103 + rans_interface.cpp:60-78
 *   - decoder incl. bypass:                       rans_interface.cpp:206-275 + rans64.h:107-142 + cpp:80-96
 *   - pmf -> 16-bit quantised CDF with zero-frequency stealing:  cpp_exts/ops/ops.cpp:10-67
 * (paths under /root/reference/src/compress or /root/reference/src/third_party).
 *
 * Pinned against the compiled reference (oracle/_ref) and the known-answer vectors of SURVEY.md §4
 * in tests/test_oracle_coder.py.
 *
 * Unlike the reference, the output buffer is sized for the 2 flush words, so 1..3 symbol streams
 * are safe here (reference bug: rans_interface.cpp:170).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16u
#define BYPASS_BITS 4u
#define BYPASS_MAX 15u
#define RANS_LOWER (1ull << 31)

typedef struct {
  uint16_t start;
  uint16_t range;
  uint8_t bypass;
} token_t;

/* number of tokens one symbol expands to (1 + escape nibbles) */
static int token_count(int32_t value, int32_t max_value) {
  uint32_t raw;
  if (value < 0)
    raw = (uint32_t)(-2 * value - 1);
  else if (value >= max_value)
    raw = (uint32_t)(2 * (value - max_value));
  else
    return 1;
  int nb = 0;
  while ((raw >> (nb * BYPASS_BITS)) != 0) ++nb;
  return 1 + nb / (int)BYPASS_MAX + 1 + nb;
}

/* returns number of bytes written, or -1 when out_cap is too small */
long rans_port_encode(const int32_t *symbols, const int32_t *indexes, long n, const int32_t *cdfs,
                      long cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, uint8_t *out,
                      long out_cap) {
  long ntok = 0;
  for (long i = 0; i < n; ++i) {
    int32_t t = indexes[i];
    ntok += token_count(symbols[i] - offsets[t], cdf_sizes[t] - 2);
  }
  token_t *tok = (token_t *)malloc(sizeof(token_t) * (size_t)(ntok > 0 ? ntok : 1));
  long k = 0;
  for (long i = 0; i < n; ++i) {
    int32_t t = indexes[i];
    const int32_t *cdf = cdfs + (long)t * cdf_stride;
    int32_t max_value = cdf_sizes[t] - 2;
    int32_t value = symbols[i] - offsets[t];
    uint32_t raw = 0;
    if (value < 0) {
      raw = (uint32_t)(-2 * value - 1);
      value = max_value;
    } else if (value >= max_value) {
      raw = (uint32_t)(2 * (value - max_value));
      value = max_value;
    }
    tok[k].start = (uint16_t)cdf[value];
    tok[k].range = (uint16_t)(cdf[value + 1] - cdf[value]);
    tok[k].bypass = 0;
    ++k;
    if (value == max_value) {
      int nb = 0;
      while ((raw >> (nb * BYPASS_BITS)) != 0) ++nb;
      int v = nb;
      while (v >= (int)BYPASS_MAX) {
        tok[k].start = BYPASS_MAX; tok[k].range = BYPASS_MAX + 1; tok[k].bypass = 1; ++k;
        v -= BYPASS_MAX;
      }
      tok[k].start = (uint16_t)v; tok[k].range = (uint16_t)(v + 1); tok[k].bypass = 1; ++k;
      for (int j = 0; j < nb; ++j) {
        uint32_t nib = (raw >> (j * BYPASS_BITS)) & BYPASS_MAX;
        tok[k].start = (uint16_t)nib; tok[k].range = (uint16_t)(nib + 1); tok[k].bypass = 1; ++k;
      }
    }
  }
  /* reverse walk; words are produced last-to-first */
  size_t cap_words = (size_t)ntok + 2;
  uint32_t *words = (uint32_t *)malloc(sizeof(uint32_t) * cap_words);
  uint32_t *ptr = words + cap_words;
  uint64_t x = RANS_LOWER;
  for (long i = ntok - 1; i >= 0; --i) {
    if (!tok[i].bypass) {
      uint32_t freq = tok[i].range;
      uint64_t x_max = ((RANS_LOWER >> PRECISION) << 32) * freq;
      if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
      x = ((x / freq) << PRECISION) + (x % freq) + tok[i].start;
    } else {
      uint32_t freq = 1u << (16 - BYPASS_BITS);
      uint64_t x_max = ((RANS_LOWER >> 16) << 32) * freq;
      if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
      x = (x << BYPASS_BITS) | tok[i].start;
    }
  }
  ptr -= 2;
  ptr[0] = (uint32_t)x;
  ptr[1] = (uint32_t)(x >> 32);
  long nbytes = (long)((words + cap_words) - ptr) * 4;
  long rc = -1;
  if (nbytes <= out_cap) {
    memcpy(out, ptr, (size_t)nbytes);
    rc = nbytes;
  }
  free(words);
  free(tok);
  return rc;
}

static inline uint32_t get_bits(uint64_t *x, const uint32_t **pp, uint32_t nbits) {
  uint32_t v = (uint32_t)(*x & ((1u << nbits) - 1));
  *x >>= nbits;
  if (*x < RANS_LOWER) { *x = (*x << 32) | **pp; ++*pp; }
  return v;
}

/* returns number of stream bytes consumed */
long rans_port_decode(const uint8_t *in, long nbytes, const int32_t *indexes, long n, const int32_t *cdfs,
                      long cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int32_t *out) {
  /* copy to aligned words (+ slack: the reference reads past short streams too) */
  size_t nwords = (size_t)(nbytes / 4) + 8;
  uint32_t *w = (uint32_t *)calloc(nwords, 4);
  memcpy(w, in, (size_t)nbytes);
  const uint32_t *p = w;
  uint64_t x = (uint64_t)p[0] | ((uint64_t)p[1] << 32);
  p += 2;
  for (long i = 0; i < n; ++i) {
    int32_t t = indexes[i];
    const int32_t *cdf = cdfs + (long)t * cdf_stride;
    int32_t size = cdf_sizes[t];
    int32_t max_value = size - 2;
    uint32_t cf = (uint32_t)(x & 0xFFFFu);
    int32_t s = 0;
    while (s < size && (uint32_t)cdf[s] <= cf) ++s; /* first entry > cf */
    s -= 1;
    uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
    x = (uint64_t)freq * (x >> PRECISION) + (x & 0xFFFFu) - start;
    if (x < RANS_LOWER) { x = (x << 32) | *p; ++p; }
    int32_t value = s;
    if (value == max_value) {
      int32_t v = (int32_t)get_bits(&x, &p, BYPASS_BITS);
      int32_t nb = v;
      while (v == (int32_t)BYPASS_MAX) { v = (int32_t)get_bits(&x, &p, BYPASS_BITS); nb += v; }
      int32_t raw = 0;
      for (int j = 0; j < nb; ++j) raw |= (int32_t)get_bits(&x, &p, BYPASS_BITS) << (j * BYPASS_BITS);
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offsets[t];
  }
  long used = (long)(p - w) * 4;
  free(w);
  return used;
}

/* cdf must hold n+1 entries. Returns 0, or -1 if no frequency can be stolen. */
int pmf_to_quantized_cdf_port(const float *pmf, int n, int precision, uint32_t *cdf) {
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
  uint32_t total = 0;
  for (int i = 0; i <= n; ++i) total += cdf[i];
  for (int i = 0; i <= n; ++i) cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
  for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      uint32_t best_freq = ~0u;
      int best = -1;
      for (int j = 0; j < n; ++j) {
        uint32_t f = cdf[j + 1] - cdf[j];
        if (f > 1 && f < best_freq) { best_freq = f; best = j; }
      }
      if (best < 0) return -1;
      if (best < i) {
        for (int j = best + 1; j <= i; ++j) cdf[j]--;
      } else {
        for (int j = i + 1; j <= best; ++j) cdf[j]++;
      }
    }
  }
  return 0;
}
