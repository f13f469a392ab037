// GPU range coder (rANS) for the quantised latents of a frame: the arithmetic-coding half of the reference's entropy
// models,
//   GaussianEncoder.encode_y / decode_y        src/models/entropy_models.py:227-341
//   BitEstimator.encode_z / decode_z           src/models/entropy_models.py:152-224
//   EntropyCoder (RansEncoder / RansDecoder)   src/models/entropy_models.py:11-81
// whose native module (MLCodec_extensions_cpp) is not part of the reference tree.  The coder restates the published
// algorithm that module is built from: the byte-wise rANS of ryg_rans (32-bit state, L = 2^23, 16-bit probabilities)
// with CompressAI's escape convention (a symbol outside the table's range is sent as the table's last entry followed by
// its magnitude in 4-bit bypass groups).  oracle/rans_oracle.py is the CPU statement of the same format; the parity
// tests require byte-identical streams in both directions.
//
// Parallel layout.  rANS is sequential per stream, so the n symbols of a tensor are cut into independent streams of
// kStreamSyms consecutive symbols; one thread codes one stream (a frame has 1.23 M y symbols = 4 800 streams).  The
// encoder writes every stream backwards into its own fixed slot, a second kernel packs the streams:
//   u32 n | u32 streams | u16 bytes[streams] | stream 0 | stream 1 | ...        (little endian)
// The decoder reads the length table, prefix-sums it per thread block and decodes every stream forwards.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dmc_b200.h"
#include "kernels.h"

namespace dmc {

constexpr int kStreamSyms = 256;
constexpr int kSlotBytes = 2560;            // worst case: 256 x (2 B symbol + 8 B of bypass groups) + 4 B of state
constexpr uint32_t kRansL = 1u << 23;
constexpr int kScaleBits = 16;
constexpr int kBypassBits = 4;

struct RansTables {
  const int32_t* cdf;       // [n_cdf][stride] cumulative frequencies, 16-bit precision (last used entry = 65536)
  const int32_t* cdf_len;   // [n_cdf] entries used (= pmf length + 2)
  const int32_t* offset;    // [n_cdf] symbol value of entry 0 (negative)
  int n_cdf, stride;
};

__device__ __forceinline__ void rans_put(uint32_t& x, uint8_t*& ptr, uint32_t start, uint32_t freq) {
  const uint32_t x_max = ((kRansL >> kScaleBits) << 8) * freq;
  while (x >= x_max) {
    *--ptr = (uint8_t)(x & 0xffu);
    x >>= 8;
  }
  x = ((x / freq) << kScaleBits) + (x % freq) + start;
}
// bypass group of kBypassBits raw bits (CompressAI's RansEncPutBits): renormalise as for a symbol of frequency
// 2^(16 - bits), then shift the bits in
__device__ __forceinline__ void rans_put_bits(uint32_t& x, uint8_t*& ptr, uint32_t val) {
  const uint32_t x_max = ((kRansL >> kScaleBits) << 8) * (1u << (kScaleBits - kBypassBits));
  while (x >= x_max) {
    *--ptr = (uint8_t)(x & 0xffu);
    x >>= 8;
  }
  x = (x << kBypassBits) | val;
}

// One thread = one stream, coded from its last symbol to its first (the decoder then reads forwards).
__global__ void k_rans_encode(RansTables t, const float* __restrict__ sym, const int32_t* __restrict__ idx, long long n,
                              uint8_t* __restrict__ slots, uint16_t* __restrict__ lens, int* __restrict__ bad) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long first = s * kStreamSyms;
  if (first >= n) return;
  const long long last = first + kStreamSyms < n ? first + kStreamSyms : n;
  uint8_t* const end = slots + (s + 1) * kSlotBytes;
  uint8_t* ptr = end;
  uint32_t x = kRansL;
  for (long long i = last - 1; i >= first; --i) {
    const int ci = idx[i];
    if (ci < 0 || ci >= t.n_cdf) { atomicOr(bad, 1); continue; }
    const int32_t* cdf = t.cdf + (long long)ci * t.stride;
    const int max_value = t.cdf_len[ci] - 2;               // the escape entry
    int value = __float2int_rn(sym[i]) - t.offset[ci];
    uint32_t raw = 0;
    bool esc = false;
    if (value < 0) {
      raw = (uint32_t)(-2 * value - 1);
      value = max_value;
      esc = true;
    } else if (value >= max_value) {
      raw = (uint32_t)(2 * (value - max_value));
      value = max_value;
      esc = true;
    }
    if (esc) {
      // forward order: [escape symbol] [group count in base-15 digits of 15, then the remainder] [groups, low first]
      int n_bypass = 0;
      while ((raw >> (n_bypass * kBypassBits)) != 0) ++n_bypass;
      for (int j = n_bypass - 1; j >= 0; --j) rans_put_bits(x, ptr, (raw >> (j * kBypassBits)) & 15u);
      int val = n_bypass;
      int full = 0;
      while (val >= 15) { val -= 15; ++full; }
      rans_put_bits(x, ptr, (uint32_t)val);
      for (int j = 0; j < full; ++j) rans_put_bits(x, ptr, 15u);
    }
    const uint32_t start = (uint32_t)cdf[value], freq = (uint32_t)cdf[value + 1] - start;
    if (freq == 0) { atomicOr(bad, 2); continue; }
    rans_put(x, ptr, start, freq);
  }
  ptr -= 4;
  ptr[0] = (uint8_t)x; ptr[1] = (uint8_t)(x >> 8); ptr[2] = (uint8_t)(x >> 16); ptr[3] = (uint8_t)(x >> 24);
  lens[s] = (uint16_t)(end - ptr);
}

// exclusive prefix sum of the stream lengths (one block; a frame has a few thousand streams)
__global__ void k_rans_offsets(const uint16_t* __restrict__ lens, long long streams, unsigned long long* __restrict__ offs,
                               unsigned long long base) {
  __shared__ unsigned long long part[1024];
  const int tid = threadIdx.x;
  const long long per = (streams + blockDim.x - 1) / blockDim.x;
  const long long a = (long long)tid * per, b = a + per < streams ? a + per : streams;
  unsigned long long sum = 0;
  for (long long i = a; i < b; ++i) sum += lens[i];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    unsigned long long run = base;
    for (int i = 0; i < (int)blockDim.x; ++i) {
      const unsigned long long v = part[i];
      part[i] = run;
      run += v;
    }
    offs[streams] = run;                                   // total size of the container
  }
  __syncthreads();
  unsigned long long run = part[tid];
  for (long long i = a; i < b; ++i) {
    offs[i] = run;
    run += lens[i];
  }
}

__global__ void k_rans_pack(const uint8_t* __restrict__ slots, const uint16_t* __restrict__ lens,
                            const unsigned long long* __restrict__ offs, long long streams, long long n,
                            uint8_t* __restrict__ out, long long cap) {
  const long long s = blockIdx.x;
  if (s >= streams) return;
  if (offs[streams] > (unsigned long long)cap) return;     // the host reports the required size
  if (s == 0 && threadIdx.x == 0) {
    const uint32_t hdr[2] = {(uint32_t)n, (uint32_t)streams};
    memcpy(out, hdr, 8);
  }
  if (threadIdx.x == 0) {
    const uint16_t l = lens[s];
    out[8 + 2 * s] = (uint8_t)l;
    out[8 + 2 * s + 1] = (uint8_t)(l >> 8);
  }
  const int len = lens[s];
  const uint8_t* src = slots + (s + 1) * kSlotBytes - len;
  uint8_t* dst = out + offs[s];
  for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ uint32_t rans_get_bits(uint32_t& x, const uint8_t*& ptr) {
  const uint32_t val = x & ((1u << kBypassBits) - 1u);
  x >>= kBypassBits;
  while (x < kRansL) x = (x << 8) | *ptr++;
  return val;
}

__global__ void k_rans_decode(RansTables t, const uint8_t* __restrict__ in, const unsigned long long* __restrict__ offs,
                              const int32_t* __restrict__ idx, long long n, float* __restrict__ out, int* __restrict__ bad) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long first = s * kStreamSyms;
  if (first >= n) return;
  const long long last = first + kStreamSyms < n ? first + kStreamSyms : n;
  const uint8_t* ptr = in + offs[s];
  const uint8_t* const stop = in + offs[s + 1];
  uint32_t x = (uint32_t)ptr[0] | ((uint32_t)ptr[1] << 8) | ((uint32_t)ptr[2] << 16) | ((uint32_t)ptr[3] << 24);
  ptr += 4;
  for (long long i = first; i < last; ++i) {
    const int ci = idx[i];
    if (ci < 0 || ci >= t.n_cdf) { atomicOr(bad, 1); out[i] = 0.f; continue; }
    const int32_t* cdf = t.cdf + (long long)ci * t.stride;
    const int max_value = t.cdf_len[ci] - 2;
    const uint32_t cum = x & 0xffffu;
    int v = 0;
    while (v < max_value && (uint32_t)cdf[v + 1] <= cum) ++v;
    const uint32_t start = (uint32_t)cdf[v], freq = (uint32_t)cdf[v + 1] - start;
    x = freq * (x >> kScaleBits) + cum - start;
    while (x < kRansL) x = (x << 8) | *ptr++;
    int value = v;
    if (v == max_value) {
      int n_bypass = 0;
      uint32_t d;
      do {
        d = rans_get_bits(x, ptr);
        n_bypass += (int)d;
      } while (d == 15u);
      uint32_t raw = 0;
      for (int j = 0; j < n_bypass; ++j) raw |= rans_get_bits(x, ptr) << (j * kBypassBits);
      value = (raw & 1u) ? -(int)((raw + 1u) >> 1) : (int)(raw >> 1) + max_value;
    }
    out[i] = (float)(value + t.offset[ci]);
  }
  if (ptr > stop) atomicOr(bad, 4);                        // a stream that ran past its end: corrupt input
}

// cdf index of every symbol of y: the scale table is log-spaced, index = nearest table entry in the log domain
// (build_index_enc / build_index_dec, src/layers/inference.py:63-84, with the clamp to [scale_min, scale_max])
__global__ void k_rans_index_gaussian(const float* __restrict__ sigma, long long n, float scale_min, float scale_max,
                                      float log_scale_min, float log_step_recip, int levels, int32_t* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = sigma[i];
  s = isnan(s) ? scale_min : fminf(fmaxf(s, scale_min), scale_max);
  int k = __float2int_rn(mul_rn(sub_rn(logf(s), log_scale_min), log_step_recip));
  idx[i] = min(max(k, 0), levels - 1);
}
// cdf index of every symbol of z (NCHW, flattened): base + channel  (BitEstimator.build_indexes, entropy_models.py:208-211)
__global__ void k_rans_index_channels(long long n, long long per_channel, int channels, int base, int32_t* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx[i] = base + (int)((i / per_channel) % channels);
}

}  // namespace dmc

using namespace dmc;

struct dmc_rans {
  RansTables t{};
  int device = 0;
  std::vector<void*> allocs;
  uint8_t* slots = nullptr;
  uint16_t* lens = nullptr;
  unsigned long long* offs = nullptr;
  int* bad = nullptr;
  long long cap_streams = 0;
  std::string error;
  ~dmc_rans() {
    for (void* p : allocs) cudaFree(p);
    release_work();
  }
  void release_work() {
    cudaFree(slots); cudaFree(lens); cudaFree(offs);
    slots = nullptr; lens = nullptr; offs = nullptr;
    cap_streams = 0;
  }
  bool reserve(long long streams) {
    if (streams <= cap_streams) return true;
    release_work();
    if (cudaMalloc(&slots, (size_t)streams * kSlotBytes) != cudaSuccess ||
        cudaMalloc(&lens, (size_t)streams * sizeof(uint16_t)) != cudaSuccess ||
        cudaMalloc(&offs, (size_t)(streams + 1) * sizeof(unsigned long long)) != cudaSuccess) {
      release_work();
      error = "rans: cudaMalloc of the stream workspace failed";
      return false;
    }
    cap_streams = streams;
    return true;
  }
};

static std::string g_rans_error;

extern "C" {

int dmc_rans_create(const int32_t* cdf, const int32_t* cdf_len, const int32_t* offset, int n_cdf, int stride,
                    dmc_rans** out) {
  if (!out) return DMC_E_INVALID;
  *out = nullptr;
  if (!cdf || !cdf_len || !offset || n_cdf < 1 || stride < 3) { g_rans_error = "dmc_rans_create: bad arguments"; return DMC_E_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_rans_error = "CUDA device required: the range coder has no CPU path";
    return DMC_E_CUDA;
  }
  for (int i = 0; i < n_cdf; ++i) {        // every table must be a strictly usable 16-bit cdf
    const int len = cdf_len[i];
    if (len < 3 || len > stride || cdf[(size_t)i * stride] != 0 || cdf[(size_t)i * stride + len - 1] != (1 << kScaleBits)) {
      g_rans_error = "dmc_rans_create: table " + std::to_string(i) + " is not a 16-bit cdf (0 ... 65536)";
      return DMC_E_INVALID;
    }
    for (int j = 0; j + 1 < len; ++j)
      if (cdf[(size_t)i * stride + j + 1] <= cdf[(size_t)i * stride + j]) {
        g_rans_error = "dmc_rans_create: table " + std::to_string(i) + " has an empty symbol";
        return DMC_E_INVALID;
      }
  }
  dmc_rans* r = new dmc_rans();
  cudaGetDevice(&r->device);
  int32_t *d_cdf = nullptr, *d_len = nullptr, *d_off = nullptr;
  if (cudaMalloc(&d_cdf, (size_t)n_cdf * stride * 4) != cudaSuccess || cudaMalloc(&d_len, (size_t)n_cdf * 4) != cudaSuccess ||
      cudaMalloc(&d_off, (size_t)n_cdf * 4) != cudaSuccess || cudaMalloc(&r->bad, 4) != cudaSuccess) {
    g_rans_error = "dmc_rans_create: cudaMalloc failed";
    delete r;
    return DMC_E_CUDA;
  }
  r->allocs = {d_cdf, d_len, d_off, r->bad};
  cudaMemcpy(d_cdf, cdf, (size_t)n_cdf * stride * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_len, cdf_len, (size_t)n_cdf * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_off, offset, (size_t)n_cdf * 4, cudaMemcpyHostToDevice);
  r->t = RansTables{d_cdf, d_len, d_off, n_cdf, stride};
  *out = r;
  return DMC_OK;
}

void dmc_rans_destroy(dmc_rans* r) { delete r; }
const char* dmc_rans_last_error(const dmc_rans* r) { return r ? r->error.c_str() : g_rans_error.c_str(); }

int dmc_rans_index_gaussian(const float* sigma, int64_t n, float scale_min, float scale_max, int levels, int32_t* idx,
                            void* stream) {
  if (n == 0) return DMC_OK;
  if (!sigma || !idx || n < 0 || levels < 2 || !(scale_min > 0.f) || !(scale_max > scale_min)) return DMC_E_INVALID;
  const float log_min = logf(scale_min);
  const float step = (logf(scale_max) - log_min) / (float)(levels - 1);
  launch(k_rans_index_gaussian, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, sigma, (long long)n, scale_min,
         scale_max, log_min, 1.0f / step, levels, idx);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

int dmc_rans_index_channels(int64_t n, int64_t per_channel, int channels, int base, int32_t* idx, void* stream) {
  if (n == 0) return DMC_OK;
  if (!idx || n < 0 || per_channel < 1 || channels < 1) return DMC_E_INVALID;
  launch(k_rans_index_channels, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, (long long)n,
         (long long)per_channel, channels, base, idx);
  return cudaGetLastError() == cudaSuccess ? DMC_OK : DMC_E_CUDA;
}

int64_t dmc_rans_max_bytes(int64_t n) {
  const int64_t streams = (n + kStreamSyms - 1) / kStreamSyms;
  return 8 + 2 * streams + streams * (int64_t)kSlotBytes;
}

// Encodes n symbols (device fp32, integer-valued) with the cdf index of each (device int32) into `out` (device, `cap`
// bytes).  Synchronises the stream; *nbytes receives the size of the container (also when it exceeds cap: DMC_E_INVALID).
int dmc_rans_encode(dmc_rans* r, const float* sym, const int32_t* idx, int64_t n, uint8_t* out, int64_t cap,
                    int64_t* nbytes, void* stream) {
  if (!r || !out || !nbytes || n < 0 || (n > 0 && (!sym || !idx))) return DMC_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const long long streams = (n + kStreamSyms - 1) / kStreamSyms;
  if (cap < 8) { r->error = "dmc_rans_encode: output buffer smaller than the header"; return DMC_E_INVALID; }
  if (streams == 0) {
    const uint32_t hdr[2] = {0u, 0u};
    cudaMemcpyAsync(out, hdr, 8, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    *nbytes = 8;
    return DMC_OK;
  }
  if (!r->reserve(streams)) return DMC_E_CUDA;
  cudaMemsetAsync(r->bad, 0, 4, st);
  launch(k_rans_encode, (unsigned)((streams + 63) / 64), 64, 0, st, r->t, sym, idx, (long long)n, r->slots, r->lens, r->bad);
  launch(k_rans_offsets, 1, 256, 0, st, (const uint16_t*)r->lens, streams, r->offs, (unsigned long long)(8 + 2 * streams));
  launch(k_rans_pack, (unsigned)streams, 64, 0, st, (const uint8_t*)r->slots, (const uint16_t*)r->lens,
         (const unsigned long long*)r->offs, streams, (long long)n, out, (long long)cap);
  unsigned long long total = 0;
  int bad = 0;
  cudaMemcpyAsync(&total, r->offs + streams, 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&bad, r->bad, 4, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { r->error = cudaGetErrorString(cudaGetLastError()); return DMC_E_CUDA; }
  *nbytes = (int64_t)total;
  if (bad) { r->error = "dmc_rans_encode: cdf index out of range or empty table entry"; return DMC_E_INVALID; }
  if ((int64_t)total > cap) { r->error = "dmc_rans_encode: output buffer too small"; return DMC_E_INVALID; }
  return DMC_OK;
}

// Decodes a container produced by dmc_rans_encode (device bytes) into n fp32 symbols; idx as for the encoder.
int dmc_rans_decode(dmc_rans* r, const uint8_t* in, int64_t nbytes, const int32_t* idx, int64_t n, float* sym_out,
                    void* stream) {
  if (!r || !in || n < 0 || nbytes < 8 || (n > 0 && (!idx || !sym_out))) return DMC_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t hdr[2];
  cudaMemcpyAsync(hdr, in, 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  const long long streams = (n + kStreamSyms - 1) / kStreamSyms;
  if ((int64_t)hdr[0] != n || (long long)hdr[1] != streams) {
    r->error = "dmc_rans_decode: container holds " + std::to_string(hdr[0]) + " symbols in " + std::to_string(hdr[1]) +
               " streams, expected " + std::to_string(n);
    return DMC_E_INVALID;
  }
  if (streams == 0) return DMC_OK;
  if (nbytes < 8 + 2 * streams) { r->error = "dmc_rans_decode: truncated length table"; return DMC_E_INVALID; }
  if (!r->reserve(streams)) return DMC_E_CUDA;
  // the u16 length table sits at byte 8 (2-byte aligned when `in` is): prefix-sum it on the device
  cudaMemcpyAsync(r->lens, in + 8, (size_t)streams * 2, cudaMemcpyDeviceToDevice, st);
  launch(k_rans_offsets, 1, 256, 0, st, (const uint16_t*)r->lens, streams, r->offs, (unsigned long long)(8 + 2 * streams));
  unsigned long long total = 0;
  cudaMemcpyAsync(&total, r->offs + streams, 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  if ((int64_t)total != nbytes) { r->error = "dmc_rans_decode: stream lengths do not add up to the container size"; return DMC_E_INVALID; }
  cudaMemsetAsync(r->bad, 0, 4, st);
  launch(k_rans_decode, (unsigned)((streams + 63) / 64), 64, 0, st, r->t, in, (const unsigned long long*)r->offs, idx,
         (long long)n, sym_out, r->bad);
  int bad = 0;
  cudaMemcpyAsync(&bad, r->bad, 4, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { r->error = cudaGetErrorString(cudaGetLastError()); return DMC_E_CUDA; }
  if (bad) { r->error = "dmc_rans_decode: corrupt stream or cdf index out of range"; return DMC_E_INVALID; }
  return DMC_OK;
}

}  // extern "C"
