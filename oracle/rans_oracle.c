/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * CPU statement (plain C) of the range coder behind the reference's EntropyCoder
 *   src/models/entropy_models.py:11-81   (RansEncoder / RansDecoder / pmf_to_quantized_cdf of MLCodec_extensions_cpp)
 * That native module is NOT in the reference tree (a pip/C++ dependency that was never vendored: the DCVC "MLCodec"
 * extension).  What is restated here is the published algorithm it is built from:
 *   - byte-wise rANS (F. Giesen, ryg_rans/rans_byte.h): 32-bit state, lower bound L = 2^23, byte renormalisation,
 *     16-bit cumulative frequencies, symbols pushed in reverse so that the decoder pops them forwards;
 *   - CompressAI's rans_interface / DCVC's rans.cpp conventions: pmf_to_quantized_cdf (round, rescale to 2^16, steal
 *     from the cheapest symbol until no entry is empty), the escape symbol (last table entry) followed by the
 *     magnitude in 4-bit bypass groups for values outside [offset, offset + max_value).
 * PARITY UNPINNED for the bit format: the reference holds no golden streams and its coder cannot be run.  What the
 * tests pin instead: round trips, byte-identical agreement between this file and the CUDA coder in both directions, the
 * table construction against the reference's own Python (GaussianEncoder.update / BitEstimator.update formulas), and
 * coded size against the ideal code length of the tables.
 *
 * Container (shared with csrc/rans.cu): u32 n | u32 streams | u16 bytes[streams] | stream ... ; a stream holds
 * RANS_STREAM consecutive symbols.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define RANS_L (1u << 23)
#define SCALE_BITS 16
#define BYPASS_BITS 4
#define RANS_STREAM 256
#define SLOT_BYTES 2560

/* CompressAI: pmf_to_quantized_cdf.  pmf: n probabilities (the last one is the tail mass); cdf: n + 1 entries. */
int rans_oracle_pmf_to_cdf(const float* pmf, int n, int precision, uint32_t* cdf) {
  uint32_t total = 0;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) {
    cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
    total += cdf[i + 1];
  }
  if (total == 0) return -1;
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
      if (best < 0) return -2;
      if (best < i) { for (int j = best + 1; j <= i; ++j) cdf[j]--; }
      else { for (int j = i + 1; j <= best; ++j) cdf[j]++; }
    }
  }
  return 0;
}

static void put(uint32_t* x, uint8_t** p, uint32_t start, uint32_t freq) {
  uint32_t x_max = ((RANS_L >> SCALE_BITS) << 8) * freq;
  while (*x >= x_max) { *--(*p) = (uint8_t)(*x & 0xff); *x >>= 8; }
  *x = ((*x / freq) << SCALE_BITS) + (*x % freq) + start;
}
static void put_bits(uint32_t* x, uint8_t** p, uint32_t val) {
  uint32_t x_max = ((RANS_L >> SCALE_BITS) << 8) * (1u << (SCALE_BITS - BYPASS_BITS));
  while (*x >= x_max) { *--(*p) = (uint8_t)(*x & 0xff); *x >>= 8; }
  *x = (*x << BYPASS_BITS) | val;
}
static uint32_t get_bits(uint32_t* x, const uint8_t** p) {
  uint32_t v = *x & ((1u << BYPASS_BITS) - 1u);
  *x >>= BYPASS_BITS;
  while (*x < RANS_L) *x = (*x << 8) | *(*p)++;
  return v;
}

/* ops of one symbol in FORWARD order; the encoder replays them backwards */
typedef struct { uint32_t start, freq; int bits; } Op;

/* Returns the container size, or -1 (bad index / empty entry / output too small). */
int64_t rans_oracle_encode(const int32_t* cdf, const int32_t* cdf_len, const int32_t* offset, int n_cdf, int stride,
                           const int32_t* sym, const int32_t* idx, int64_t n, uint8_t* out, int64_t cap) {
  int64_t streams = (n + RANS_STREAM - 1) / RANS_STREAM;
  int64_t pos = 8 + 2 * streams;
  if (cap < pos) return -1;
  uint32_t hdr[2] = {(uint32_t)n, (uint32_t)streams};
  memcpy(out, hdr, 8);
  uint8_t* slot = (uint8_t*)malloc(SLOT_BYTES);
  Op* ops = (Op*)malloc(sizeof(Op) * RANS_STREAM * 16);
  for (int64_t s = 0; s < streams; ++s) {
    int64_t first = s * RANS_STREAM, last = first + RANS_STREAM < n ? first + RANS_STREAM : n;
    int nops = 0;
    for (int64_t i = first; i < last; ++i) {
      int ci = idx[i];
      if (ci < 0 || ci >= n_cdf) { free(slot); free(ops); return -1; }
      const int32_t* c = cdf + (int64_t)ci * stride;
      int max_value = cdf_len[ci] - 2;
      int value = sym[i] - offset[ci];
      uint32_t raw = 0;
      int esc = 0;
      if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; esc = 1; }
      else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; esc = 1; }
      uint32_t start = (uint32_t)c[value], freq = (uint32_t)c[value + 1] - start;
      if (freq == 0) { free(slot); free(ops); return -1; }
      ops[nops++] = (Op){start, freq, 0};
      if (esc) {
        int n_bypass = 0;
        while ((raw >> (n_bypass * BYPASS_BITS)) != 0) ++n_bypass;
        int val = n_bypass;
        while (val >= 15) { ops[nops++] = (Op){15u, 0, 1}; val -= 15; }
        ops[nops++] = (Op){(uint32_t)val, 0, 1};
        for (int j = 0; j < n_bypass; ++j) ops[nops++] = (Op){(raw >> (j * BYPASS_BITS)) & 15u, 0, 1};
      }
    }
    uint8_t* p = slot + SLOT_BYTES;
    uint32_t x = RANS_L;
    for (int k = nops - 1; k >= 0; --k) {
      if (ops[k].bits) put_bits(&x, &p, ops[k].start);
      else put(&x, &p, ops[k].start, ops[k].freq);
    }
    p -= 4;
    p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
    int len = (int)(slot + SLOT_BYTES - p);
    if (pos + len > cap) { free(slot); free(ops); return -1; }
    out[8 + 2 * s] = (uint8_t)len;
    out[8 + 2 * s + 1] = (uint8_t)(len >> 8);
    memcpy(out + pos, p, len);
    pos += len;
  }
  free(slot);
  free(ops);
  return pos;
}

/* Returns 0, or a negative code for a malformed container. */
int rans_oracle_decode(const int32_t* cdf, const int32_t* cdf_len, const int32_t* offset, int n_cdf, int stride,
                       const uint8_t* in, int64_t nbytes, const int32_t* idx, int64_t n, int32_t* sym) {
  if (nbytes < 8) return -1;
  uint32_t hdr[2];
  memcpy(hdr, in, 8);
  int64_t streams = (n + RANS_STREAM - 1) / RANS_STREAM;
  if ((int64_t)hdr[0] != n || (int64_t)hdr[1] != streams) return -2;
  if (nbytes < 8 + 2 * streams) return -3;
  int64_t pos = 8 + 2 * streams;
  for (int64_t s = 0; s < streams; ++s) {
    int len = in[8 + 2 * s] | (in[8 + 2 * s + 1] << 8);
    if (pos + len > nbytes || len < 4) return -4;
    const uint8_t* p = in + pos;
    const uint8_t* stop = p + len;
    uint32_t x = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    p += 4;
    int64_t first = s * RANS_STREAM, last = first + RANS_STREAM < n ? first + RANS_STREAM : n;
    for (int64_t i = first; i < last; ++i) {
      int ci = idx[i];
      if (ci < 0 || ci >= n_cdf) return -5;
      const int32_t* c = cdf + (int64_t)ci * stride;
      int max_value = cdf_len[ci] - 2;
      uint32_t cum = x & 0xffffu;
      int v = 0;
      while (v < max_value && (uint32_t)c[v + 1] <= cum) ++v;
      uint32_t start = (uint32_t)c[v], freq = (uint32_t)c[v + 1] - start;
      x = freq * (x >> SCALE_BITS) + cum - start;
      while (x < RANS_L) x = (x << 8) | *p++;
      int value = v;
      if (v == max_value) {
        int n_bypass = 0;
        uint32_t d;
        do { d = get_bits(&x, &p); n_bypass += (int)d; } while (d == 15u);
        uint32_t raw = 0;
        for (int j = 0; j < n_bypass; ++j) raw |= get_bits(&x, &p) << (j * BYPASS_BITS);
        value = (raw & 1u) ? -(int)((raw + 1u) >> 1) : (int)(raw >> 1) + max_value;
      }
      sym[i] = value + offset[ci];
    }
    if (p > stop) return -6;
    pos += len;
  }
  return pos == nbytes ? 0 : -7;
}
