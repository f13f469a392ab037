// Training-mode kernels: what the backward pass of a DepthConvBlock (layers.py:43-79) and of the
// quantisation / likelihood terms (inference.py:16-27, entropy_models.py:303-341) needs beyond the forward kernels.
//
//   * data gradients of the 1x1 convolutions are ordinary contractions with the transposed weight: they run on the
//     forward tcgen05 chain kernel (gemm_s3.cu) at the same 3-term split-fp16 accuracy;
//   * weight gradients dW[n][k] = sum_m G[m][n] X[m][k] contract over the PIXEL axis, i.e. both operands are read
//     "MN-major" from the pixel-major split planes.  The product kernel is k_wgrad_umma (wgrad_umma.cu: tcgen05 with
//     MN-major descriptors); k_wgrad_s3 here does the same with mma.sync m16n8k16 + ldmatrix.trans straight from the
//     tile-blocked planes (their 32-byte swizzle is conflict-free for ldmatrix as it is) and stays as the A/B and
//     validation kernel (DMC_WGRAD_UMMA=0: the legacy tensor path measured 110 T MAC/s, 1/16 of tcgen05's peak).
//     Both split the pixel axis into per-CTA partial sums that k_reduce_partials adds up in a fixed order;
//   * everything elementwise (WSiLU', chunk-add', depthwise weight gradient, column sums for the biases) is
//     HBM-bound and vectorised 8 channels per thread on the split planes.
//
// Gradients are linear in the incoming gradient: it is multiplied by a power of two found on the device (k_absmax_part,
// k_grad_scale: max |g| -> 2^8) on its way into the fp16 split planes, and every result leaves multiplied by the
// reciprocal (`scale_dev` of the reductions, k_s3_to_nchw_scaled).  The launch order lives in engine.cu
// (dmc_dcb_train_create).
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

#include "kernels.h"

namespace dmc {

static inline unsigned cdiv_u(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Opt a kernel in to the device's largest dynamic shared-memory size, once per (kernel, device): the launchers below are
// also run while a CUDA graph is being captured, where attribute calls are best not made at all.
template <class K>
static void optin_max_smem(K kernel) {
  static std::set<std::pair<const void*, int>> done;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if (!done.insert({(const void*)kernel, dev}).second) return;
  int smem = 0;
  cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// sigmoid with the hardware exponential and reciprocal (2 ulp): only used where the result feeds a gradient -- the
// forward values the reference pins bit by bit never pass through here
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
// d/dx [ silu(4x) / 4 ] = sigmoid(v) * (1 + v * (1 - sigmoid(v))),  v = 4x          (layers.py:8-10)
__device__ __forceinline__ float wsilu_grad(float x) {
  const float v = 4.0f * x;
  const float s = sigmoid_fast(v);
  return s * fmaf(v, 1.0f - s, 1.0f);
}

// ------------------------------------------------------------------ elementwise, fp32 rows / split planes
// out = wsilu(in), fp32 rows [M, C]
__global__ void k_wsilu_rows(const float* __restrict__ in, float* __restrict__ out, long long n4) {
  pdl_prologue_done();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<const float4*>(in)[i];
  a.x = wsilu(a.x); a.y = wsilu(a.y); a.z = wsilu(a.z); a.w = wsilu(a.w);
  reinterpret_cast<float4*>(out)[i] = a;
}
void wsilu_rows(const float* in, float* out, long long n, cudaStream_t st) {
  launch(k_wsilu_rows, cdiv_u(n / 4, 256), 256, 0, st, in, out, n / 4);
}

// out[m, c] = g[m, c] * wsilu'(pre[m, c]);   g, out: split planes [M, C];  pre: fp32 rows [M, ld];
// part[block][c] = the block's column sums of out (dc.0's bias gradient).  Persistent grid, 8 columns per thread.
__global__ void __launch_bounds__(256) k_wsilu_bwd(View g, const float* __restrict__ pre, int ld, View out, long long M, int C8,
                                                   float* __restrict__ part, int ldp) {
  pdl_prologue_done();
  extern __shared__ float red[];        // [lanes][C8 * 8]
  const int lanes = blockDim.x / C8;
  const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
  const int c = c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    for (long long m = (long long)blockIdx.x * lanes + rl; m < M; m += (long long)gridDim.x * lanes) {
      float v[8];
      ld3x8(g, m, c, v);
      const float4 p0 = *reinterpret_cast<const float4*>(pre + m * ld + c);
      const float4 p1 = *reinterpret_cast<const float4*>(pre + m * ld + c + 4);
      const float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] *= wsilu_grad(p[i]);
        acc[i] += v[i];
      }
      st3x8(out, m, c, v);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(rl * C8 + c8) * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C8 * 8; j += blockDim.x) {
    float s = 0.0f;
    for (int r = 0; r < lanes; ++r) s += red[r * C8 * 8 + j];
    part[(long long)blockIdx.x * ldp + j] = s;
  }
}
// the number of partial rows wsilu_bwd writes for these sizes
int wsilu_bwd_parts(long long M, int C, int max_parts) {
  const int C8 = C / 8;
  int threads = 256;
  if (C8 > threads) threads = (C8 + 31) / 32 * 32;
  const int lanes = threads / C8;
  int blocks = (int)((M + lanes - 1) / lanes);
  const int want = 4 * num_sms();
  if (blocks > want) blocks = want;
  if (blocks > max_parts) blocks = max_parts;
  if (blocks < 1) blocks = 1;
  return blocks;
}
int wsilu_bwd(View g, const float* pre, int ld, View out, long long M, float* part, int ldp, int max_parts, cudaStream_t st) {
  const int C8 = g.C / 8;
  int threads = 256;
  if (C8 > threads) threads = (C8 + 31) / 32 * 32;
  const int lanes = threads / C8;
  const int blocks = wsilu_bwd_parts(M, g.C, max_parts);
  const size_t smem = (size_t)lanes * C8 * 8 * sizeof(float);
  launch(k_wsilu_bwd, blocks, threads, smem, st, g, pre, ld, out, M, C8, part, ldp);
  return blocks;
}

// The two directions of WSiLUChunkAdd (layers.py:12-20) in one pass over the pre-activations u [M, 2*C2] (fp32 rows):
//   v[m, j]  = wsilu(u[m, j]) + wsilu(u[m, j + C2])              (the recomputed forward value, operand of ffn.2's dW)
//   gu[m, j] = gv[m, j mod C2] * wsilu'(u[m, j])                 (gradient of the pre-activations)
//   part[block][j] = sum over the block's rows of gu[m, j]       (ffn.0's bias gradient, reduced by k_reduce_partials)
// A thread owns 8 columns (of both halves) and kCaRows rows; a block's partial row covers lanes * kCaRows rows (one
// pass of ~1 200 blocks at 160 x 240 x 512: a persistent grid of 4 blocks per SM ran this kernel at 3.2 TB/s, latency bound).
constexpr int kCaRows = 8;
__global__ void __launch_bounds__(256) k_chunkadd_fwd_bwd(const float* __restrict__ u, int ld, View gv, View v, View gu,
                                                          long long M, int C2, float* __restrict__ part, int ldp) {
  pdl_prologue_done();
  extern __shared__ float red[];        // [lanes][2 * C2]
  const int C8 = C2 / 8;
  const int lanes = blockDim.x / C8;
  const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
  const int c = c8 * 8;
  float acc[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[h][i] = 0.0f;
  if (rl < lanes) {
    const long long m0 = (long long)blockIdx.x * lanes * kCaRows + rl;
#pragma unroll 2
    for (int it = 0; it < kCaRows; ++it) {
      const long long m = m0 + (long long)it * lanes;
      if (m >= M) break;
      float g[8], o[8];
      ld3x8(gv, m, c, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.0f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float* q = u + m * ld + half * C2 + c;
        const float4 p0 = *reinterpret_cast<const float4*>(q);
        const float4 p1 = *reinterpret_cast<const float4*>(q + 4);
        const float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // one exponential for the value and the derivative: s = sigmoid(4p); wsilu = p * s; wsilu' = s * (1 + 4p(1 - s))
          const float t = 4.0f * p[i];
          const float sg = sigmoid_fast(t);
          o[i] += p[i] * sg;
          d[i] = g[i] * (sg * fmaf(t, 1.0f - sg, 1.0f));
          acc[half][i] += d[i];
        }
        st3x8(gu, m, half * C2 + c, d);
      }
      st3x8(v, m, c, o);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[(long long)rl * 2 * C2 + h * C2 + c + i] = acc[h][i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * C2; j += blockDim.x) {
    float s = 0.0f;
    for (int r = 0; r < lanes; ++r) s += red[(long long)r * 2 * C2 + j];
    part[(long long)blockIdx.x * ldp + j] = s;
  }
}
// returns the number of partial rows (2 * C2 floats each, pitch ldp)
int chunkadd_parts(long long M, int C2) {
  const int C8 = C2 / 8;
  int threads = 256;
  if (C8 > threads) threads = (C8 + 31) / 32 * 32;
  const long long per_block = (long long)(threads / C8) * kCaRows;
  return (int)((M + per_block - 1) / per_block);
}
int chunkadd_fwd_bwd(const float* u, int ld, View gv, View v, View gu, long long M, float* part, int ldp, int max_parts,
                     cudaStream_t st) {
  const int C2 = gv.C, C8 = C2 / 8;
  int threads = 256;
  if (C8 > threads) threads = (C8 + 31) / 32 * 32;
  const int lanes = threads / C8;
  const long long per_block = (long long)lanes * kCaRows;
  const int blocks = (int)((M + per_block - 1) / per_block);
  if (blocks > max_parts) return -1;
  const size_t smem = (size_t)lanes * 2 * C2 * sizeof(float);
  optin_max_smem(k_chunkadd_fwd_bwd);
  launch(k_chunkadd_fwd_bwd, blocks, threads, smem, st, u, ld, gv, v, gu, M, C2, part, ldp);
  return blocks;
}

// ------------------------------------------------------------------ column sums (bias gradients, per-channel scale)
// part[blockIdx.x][c] = sum over this block's rows of g[m, c] (* h[m, c] when h is given)
__global__ void __launch_bounds__(256) k_colsum_s3(View g, View h, int has_h, long long M, int C8, float* __restrict__ part,
                                                   int ldp) {
  pdl_prologue_done();
  extern __shared__ float red[];        // [rows_per_block][C8 * 8]
  const int lanes = blockDim.x / C8;    // row lanes per block
  const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    for (long long m = (long long)blockIdx.x * lanes + rl; m < M; m += (long long)gridDim.x * lanes) {
      float v[8];
      ld3x8(g, m, c8 * 8, v);
      if (has_h) {
        float w[8];
        ld3x8(h, m, c8 * 8, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], w[i], acc[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(rl * C8 + c8) * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += blockDim.x) {
    float s = 0.0f;
    for (int r = 0; r < lanes; ++r) s += red[r * C8 * 8 + c];
    part[(long long)blockIdx.x * ldp + c] = s;
  }
}
// returns the number of partial rows written to `part` (each `ldp` floats apart)
int colsum_s3(View g, const View* h, long long M, float* part, int ldp, int max_parts, cudaStream_t st) {
  const int C8 = (g.C + 7) / 8;
  int threads = 256;
  if (C8 > threads) threads = (C8 + 31) / 32 * 32;
  const int lanes = threads / C8;
  int blocks = (int)((M + lanes - 1) / lanes);
  const int want = 8 * num_sms();      // (2 per SM: 2.9 TB/s at 24 % occupancy; more rows in flight)
  if (blocks > want) blocks = want;
  if (blocks > max_parts) blocks = max_parts;
  if (blocks < 1) blocks = 1;
  View hv = h ? *h : g;
  launch(k_colsum_s3, blocks, threads, (size_t)lanes * C8 * 8 * sizeof(float), st, g, hv, h ? 1 : 0, M, C8, part, ldp);
  return blocks;
}

// out[i] = scale * sum_s part[s * stride + i]   (fixed summation order: the result does not depend on scheduling)
__global__ void k_reduce_partials(const float* __restrict__ part, long long stride, int S, float* __restrict__ out,
                                  long long n, const float* __restrict__ scale_dev, float scale) {
  pdl_prologue_done();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // four independent partial sums keep four loads in flight (the loop is latency bound otherwise); the order of the
  // additions is fixed by S alone
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  int k = 0;
  for (; k + 3 < S; k += 4) {
    s0 += part[(long long)k * stride + i];
    s1 += part[(long long)(k + 1) * stride + i];
    s2 += part[(long long)(k + 2) * stride + i];
    s3 += part[(long long)(k + 3) * stride + i];
  }
  for (; k < S; ++k) s0 += part[(long long)k * stride + i];
  if (scale_dev) scale *= *scale_dev;
  out[i] = ((s0 + s1) + (s2 + s3)) * scale;
}
// many partial rows, few outputs (column sums): 32 outputs x 32 row lanes per block, four loads in flight per thread, the
// lanes combined in a fixed order
__global__ void __launch_bounds__(1024) k_reduce_partials_tall(const float* __restrict__ part, long long stride, int S,
                                                               float* __restrict__ out, long long n,
                                                               const float* __restrict__ scale_dev, float scale) {
  pdl_prologue_done();
  __shared__ float red[32][33];
  const int col = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + col;
  float s = 0.0f;
  if (i < n) {
    float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int k = lane;
    for (; k + 96 < S; k += 128) {
      s += part[(long long)k * stride + i];
      s1 += part[(long long)(k + 32) * stride + i];
      s2 += part[(long long)(k + 64) * stride + i];
      s3 += part[(long long)(k + 96) * stride + i];
    }
    for (; k < S; k += 32) s += part[(long long)k * stride + i];
    s = (s + s1) + (s2 + s3);
  }
  red[lane][col] = s;
  __syncthreads();
  if (lane == 0 && i < n) {
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += red[r][col];
    if (scale_dev) scale *= *scale_dev;
    out[i] = t * scale;
  }
}
void reduce_partials(const float* part, long long stride, int S, float* out, long long n, const float* scale_dev,
                     float scale, cudaStream_t st) {
  if (S >= 32 && n <= 8192) launch(k_reduce_partials_tall, cdiv_u(n, 32), 1024, 0, st, part, stride, S, out, n, scale_dev, scale);
  else launch(k_reduce_partials, cdiv_u(n, 256), 256, 0, st, part, stride, S, out, n, scale_dev, scale);
}

// ------------------------------------------------------------------ depthwise 3x3: weight / bias gradient
// dW[c][ky][kx] = sum_{b,h,w} g[b,h,w,c] * t[b,h+ky-1,w+kx-1,c]  (zero outside), db[c] = sum g.
// g: fp32 rows [M, ldg] (what dc.3's data gradient writes); t: fp32 rows [M, ld] (the WSiLU output the forward conv read).
// Thread = 8 channels x one pixel lane; part[block][c*10 + tap] (tap 9 = bias).
__global__ void __launch_bounds__(256) k_dw_wgrad(const float* __restrict__ g, int ldg, const float* __restrict__ t, int ld, int B,
                                                  int H, int W, int C8, float* __restrict__ part, int ldp) {
  pdl_prologue_done();
  extern __shared__ float red[];        // [lanes][C8*8*10]
  const int lanes = blockDim.x / C8;
  const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
  const long long M = (long long)B * H * W;
  float acc[10][8];
#pragma unroll
  for (int k = 0; k < 10; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.0f;
  if (rl < lanes) {
    for (long long m = (long long)blockIdx.x * lanes + rl; m < M; m += (long long)gridDim.x * lanes) {
      const int w = (int)(m % W);
      const int h = (int)((m / W) % H);
      const float4 g0 = *reinterpret_cast<const float4*>(g + m * ldg + c8 * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(g + m * ldg + c8 * 8 + 4);
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[9][i] += gv[i];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int hh = h + ky - 1;
        if ((unsigned)hh >= (unsigned)H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ww = w + kx - 1;
          if ((unsigned)ww >= (unsigned)W) continue;
          const float* q = t + (m + (long long)(ky - 1) * W + (kx - 1)) * ld + c8 * 8;
          const float4 a = *reinterpret_cast<const float4*>(q);
          const float4 b = *reinterpret_cast<const float4*>(q + 4);
          const float tv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[ky * 3 + kx][i] = fmaf(gv[i], tv[i], acc[ky * 3 + kx][i]);
        }
      }
    }
    float* r = red + (long long)rl * C8 * 80;
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) r[(c8 * 8 + i) * 10 + k] = acc[k][i];
  }
  __syncthreads();
  const int n = C8 * 80;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float s = 0.0f;
    for (int r = 0; r < lanes; ++r) s += red[(long long)r * n + j];
    part[(long long)blockIdx.x * ldp + j] = s;
  }
}
// Strip version: a thread owns 4 channels of ONE image column and walks down a strip of rows with the 3 x 3 window of t
// in registers (three new float4 loads + one of g per pixel instead of nine + one: the one-pixel-per-iteration kernel
// above re-reads every t row three times and ran at 1.3 TB/s of algorithmic bytes, 20 % of the HBM roofline).
__global__ void __launch_bounds__(256, 2)
k_dw_wgrad_strip(const float* __restrict__ g, int ldg, const float* __restrict__ t, int ld, int B, int H, int W, int C4, int lanes,
                 int WG, int hs, int HS, float* __restrict__ part, int ldp) {
  pdl_prologue_done();
  extern __shared__ float red[];        // [lanes][C4 * 4 * 10]
  const int c4 = threadIdx.x % C4, ln = threadIdx.x / C4;
  int blk = blockIdx.x;
  const int wg = blk % WG;
  blk /= WG;
  const int hsi = blk % HS, b = blk / HS;
  const int w = wg * lanes + ln;
  const int h0 = hsi * hs, h1 = min(H, h0 + hs);
  float acc[9][4], accb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f;
  if (ln < lanes && w < W) {
    const int c = c4 * 4;
    const long long img = (long long)b * H * W;
    float4 win[3][3];
    auto load_row = [&](int h, float4 (&r)[3]) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        r[kx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if ((unsigned)h < (unsigned)H && (unsigned)ww < (unsigned)W)
          r[kx] = *reinterpret_cast<const float4*>(t + (img + (long long)h * W + ww) * ld + c);
      }
    };
    load_row(h0 - 1, win[0]);
    load_row(h0, win[1]);
    for (int h = h0; h < h1; ++h) {
      load_row(h + 1, win[2]);
      const float4 gv = *reinterpret_cast<const float4*>(g + (img + (long long)h * W + w) * ldg + c);
      accb[0] += gv.x; accb[1] += gv.y; accb[2] += gv.z; accb[3] += gv.w;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 tv = win[ky][kx];
          float* a = acc[ky * 3 + kx];
          a[0] = fmaf(gv.x, tv.x, a[0]); a[1] = fmaf(gv.y, tv.y, a[1]);
          a[2] = fmaf(gv.z, tv.z, a[2]); a[3] = fmaf(gv.w, tv.w, a[3]);
        }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) { win[0][kx] = win[1][kx]; win[1][kx] = win[2][kx]; }
    }
  }
  const int n = C4 * 40;
  if (ln < lanes) {
    float* r = red + (long long)ln * n + c4 * 40;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int k = 0; k < 9; ++k) r[i * 10 + k] = acc[k][i];
      r[i * 10 + 9] = accb[i];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float sacc = 0.0f;
    for (int r = 0; r < lanes; ++r) sacc += red[(long long)r * n + j];
    part[(long long)blockIdx.x * ldp + j] = sacc;
  }
}
int dw_wgrad(const float* g, int ldg, int C, const float* t, int ld, int B, int H, int W, float* part, int ldp, int max_parts,
             cudaStream_t st) {
  static int strip = -1;
  if (strip < 0) {
    const char* v = getenv("DMC_DW_WGRAD_STRIP");       // =0: the one-pixel-per-iteration kernel (A/B runs)
    strip = (v && v[0] == '0') ? 0 : 1;
  }
  if (strip && C % 4 == 0 && ld % 4 == 0 && ldg % 4 == 0 && C / 4 <= 256) {
    const int C4 = C / 4;
    const int lanes = std::max(1, 256 / C4);
    const int WG = (W + lanes - 1) / lanes;
    int HS = (4 * num_sms() + B * WG - 1) / (B * WG);        // strips per image column group: ~4 blocks per SM
    if (HS < 1) HS = 1;
    if (HS > H) HS = H;
    if ((long long)B * WG * HS > max_parts) HS = std::max(1, max_parts / (B * WG));
    const int hs = (H + HS - 1) / HS;
    HS = (H + hs - 1) / hs;
    const long long blocks = (long long)B * WG * HS;
    if (blocks <= max_parts) {
      const size_t smem = (size_t)lanes * C4 * 40 * sizeof(float);
      optin_max_smem(k_dw_wgrad_strip);
      launch(k_dw_wgrad_strip, (unsigned)blocks, C4 * lanes, smem, st, g, ldg, t, ld, B, H, W, C4, lanes, WG, hs, HS, part, ldp);
      return (int)blocks;
    }
  }

  const int C8 = C / 8;
  int threads = 256;
  int lanes = threads / C8;
  if (lanes < 1) { lanes = 1; threads = (C8 + 31) / 32 * 32; }
  // shared memory: lanes * C * 10 floats (256 channels x 8 lanes = 80 KB): opt in once
  const size_t smem = (size_t)lanes * C8 * 80 * sizeof(float);
  optin_max_smem(k_dw_wgrad);
  const long long M = (long long)B * H * W;
  int blocks = (int)((M + lanes - 1) / lanes);
  const int want = 2 * num_sms();
  if (blocks > want) blocks = want;
  if (blocks > max_parts) blocks = max_parts;
  if (blocks < 1) blocks = 1;
  launch(k_dw_wgrad, blocks, threads, smem, st, g, ldg, t, ld, B, H, W, C8, part, ldp);
  return blocks;
}


// w9c [9][C] -> the taps of the data gradient: out[tap][c] = w9c[8 - tap][c] (the kernel rotated by 180 degrees)
__global__ void k_flip_dw(const float* __restrict__ w9c, float* __restrict__ out, int C) {
  pdl_prologue_done();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  out[i] = w9c[(8 - i / C) * C + i % C];
}
void flip_dw_weight(const float* w9c, float* out, int C, cudaStream_t st) { launch(k_flip_dw, cdiv_u(9 * C, 256), 256, 0, st, w9c, out, C); }

// partial rows of dw_wgrad -> weight gradient (C,1,3,3) and bias gradient (C); either destination may be null
__global__ void __launch_bounds__(256) k_reduce_dw(const float* __restrict__ part, long long stride, int S, float* gw, float* gb,
                                                   int C, const float* __restrict__ scale_dev, float scale) {
  pdl_prologue_done();
  if (scale_dev) scale *= *scale_dev;
  __shared__ float red[8][33];
  const int col = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  float s = 0.0f;
  if (i < C * 10)
    for (int k = lane; k < S; k += 8) s += part[(long long)k * stride + i];
  red[lane][col] = s;
  __syncthreads();
  if (lane != 0 || i >= C * 10) return;
  float t = 0.0f;
#pragma unroll
  for (int r = 0; r < 8; ++r) t += red[r][col];
  const int c = i / 10, tap = i % 10;
  if (tap < 9) { if (gw) gw[c * 9 + tap] = t * scale; }
  else if (gb) gb[c] = t * scale;
}
void reduce_dw(const float* part, long long stride, int S, float* gw, float* gb, int C, const float* scale_dev, float scale,
               cudaStream_t st) {
  launch(k_reduce_dw, cdiv_u(C * 10, 32), 256, 0, st, part, stride, S, gw, gb, C, scale_dev, scale);
}

// ------------------------------------------------------------------ weight gradient of a 1x1 convolution
// part[z][n][k] = sum over the pixel slice z of G[m][n] * X[m][k].
//
// CTA = 128 (n) x 64 (k) outputs, 8 warps of 32 x 32, pixel chunks of 32 through a 3-deep cp.async ring.  A chunk of a
// plane is one contiguous 1 KB piece per 16-column block -- [32 pixels][16 columns] with the two 16-byte halves of a
// row swapped where bit 2 of the pixel index is set -- and is copied verbatim: the eight rows an ldmatrix reads (one
// 16-byte unit each, consecutive pixels, same column group) then fall into eight different 16-byte bank groups.
// Both fragments come from ldmatrix.trans (the contraction index is the ROW of the stored tiles).  Split product as
// in the forward kernels: main += Gh.Xh, small += Gh.Xl + Gl.Xh (2^11-scaled), result = main + small * 2^-11.
constexpr int kWgTN = 128, kWgTK = 64, kWgChunk = 32, kWgStages = 3;
constexpr int kWgGBytes = (kWgTN / 16) * 1024;       // one plane of the G tile of a chunk
constexpr int kWgXBytes = (kWgTK / 16) * 1024;
constexpr int kWgStageBytes = 2 * (kWgGBytes + kWgXBytes);

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of the 16-byte unit (pixel r of the chunk, column group cg of 8 columns) inside a plane tile
__device__ __forceinline__ uint32_t wg_unit(int r, int cg) {
  return (uint32_t)((cg >> 1) * 1024 + r * 32 + ((((cg & 1) ^ ((r >> 2) & 1))) << 4));
}

template <int kTerms>
__global__ void __launch_bounds__(256, 2)
k_wgrad_s3(View G, View X, int N, int K, int chunks_total, int chunks_per_split, float* __restrict__ part) {
  pdl_prologue_done();
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * kWgTN, k0 = blockIdx.y * kWgTK;
  const int c_begin = blockIdx.z * chunks_per_split;
  int c_end = c_begin + chunks_per_split;
  if (c_end > chunks_total) c_end = chunks_total;
  const int nch = c_end - c_begin;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);

  // column blocks past the tensor stay zero for the whole kernel (ragged N / K)
  for (int i = tid; i < kWgStages * kWgStageBytes / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const int gblk = (G.C + 15) / 16, xblk = (X.C + 15) / 16;
  constexpr int kPl = (kTerms == 1) ? 1 : 2;
  // a stage = [G hi][G lo][X hi][X lo]; 16-byte pieces: G 512 per plane, X 256 per plane
  auto issue = [&](int chunk, int stage) {
    const uint32_t s0 = sbase + stage * kWgStageBytes;
    const long long row0 = (long long)chunk * kWgChunk;
#pragma unroll
    for (int it = 0; it < (kPl * (kWgGBytes + kWgXBytes) / 16 + 255) / 256; ++it) {
      const int p = it * 256 + tid;
      if (p >= kPl * (kWgGBytes + kWgXBytes) / 16) break;
      const int per_pl = (kWgGBytes + kWgXBytes) / 16;      // 768
      const int pl = p / per_pl, q = p % per_pl;
      if (q < kWgGBytes / 16) {
        const int blk = q >> 6, u = q & 63;                  // 64 units per 1 KB block
        const int gb = (n0 >> 4) + blk;
        if (gb < gblk)
          cp_async16(s0 + pl * kWgGBytes + blk * 1024 + u * 16,
                     G.p + pl * G.ps + (long long)gb * G.bs + row0 * 16 + u * 8);
      } else {
        const int qq = q - kWgGBytes / 16;
        const int blk = qq >> 6, u = qq & 63;
        const int xb = (k0 >> 4) + blk;
        if (xb < xblk)
          cp_async16(s0 + 2 * kWgGBytes + pl * kWgXBytes + blk * 1024 + u * 16,
                     X.p + pl * X.ps + (long long)xb * X.bs + row0 * 16 + u * 8);
      }
    }
  };

  float acc[2][4][4], acs[(kTerms == 1) ? 1 : 2][(kTerms == 1) ? 1 : 4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[i][j][e] = 0.0f;
        if constexpr (kTerms != 1) acs[i][j][e] = 0.0f;
      }

  const int wn = (warp & 3) * 32, wk = (warp >> 2) * 32;     // warp tile origin inside the CTA tile
  // ldmatrix lane roles.  A (from G, rows = pixels): matrix j = lane >> 3: pixels (j >> 1) * 8 + (lane & 7), columns
  // (j & 1) * 8.  B (from X): pixels (j & 1) * 8 + (lane & 7), columns (j >> 1) * 8.
  const int lj = lane >> 3, lr = lane & 7;
  const int a_r = (lj >> 1) * 8 + lr, a_c = (lj & 1) * 8;
  const int b_r = (lj & 1) * 8 + lr, b_c = (lj >> 1) * 8;

  for (int s = 0; s < kWgStages - 1; ++s) {
    if (s < nch) issue(c_begin + s, s);
    cp_async_commit();
  }
  for (int i = 0; i < nch; ++i) {
    cp_async_wait<kWgStages - 2>();
    __syncthreads();
    if (i + kWgStages - 1 < nch) issue(c_begin + i + kWgStages - 1, (i + kWgStages - 1) % kWgStages);
    cp_async_commit();
    const uint32_t s0 = sbase + (i % kWgStages) * kWgStageBytes;
    const uint32_t gh = s0, gl = s0 + kWgGBytes, xh = s0 + 2 * kWgGBytes, xl = xh + kWgXBytes;
#pragma unroll
    for (int ms = 0; ms < kWgChunk / 16; ++ms) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int i16 = 0; i16 < 2; ++i16) {
        const uint32_t off = wg_unit(ms * 16 + a_r, (wn + i16 * 16 + a_c) >> 3);
        ldsm_x4_t(gh + off, ah[i16]);
        if constexpr (kTerms != 1) ldsm_x4_t(gl + off, al[i16]);
      }
#pragma unroll
      for (int j16 = 0; j16 < 2; ++j16) {
        uint32_t bh[4], bl[4];
        const uint32_t off = wg_unit(ms * 16 + b_r, (wk + j16 * 16 + b_c) >> 3);
        ldsm_x4_t(xh + off, bh);
        if constexpr (kTerms != 1) ldsm_x4_t(xl + off, bl);
#pragma unroll
        for (int i16 = 0; i16 < 2; ++i16)
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            mma16816(acc[i16][j16 * 2 + j8], ah[i16], bh[2 * j8], bh[2 * j8 + 1]);
            if constexpr (kTerms != 1) {
              mma16816(acs[i16][j16 * 2 + j8], ah[i16], bl[2 * j8], bl[2 * j8 + 1]);
              mma16816(acs[i16][j16 * 2 + j8], al[i16], bh[2 * j8], bh[2 * j8 + 1]);
            }
          }
      }
    }
  }
  cp_async_wait<0>();

  float* dst = part + (long long)blockIdx.z * N * K;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i16 = 0; i16 < 2; ++i16)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + wn + i16 * 16 + g + ((e >> 1) << 3);
        const int k = k0 + wk + j * 8 + 2 * t + (e & 1);
        if (n < N && k < K) {
          float v = acc[i16][j][e];
          if constexpr (kTerms != 1) v = fmaf(acs[i16][j][e], kLoInv, v);
          dst[(long long)n * K + k] = v;
        }
      }
}

static int wgrad_mma_splits(long long M, int N, int K) {
  const int chunks = (int)((M + kWgChunk - 1) / kWgChunk);
  const int tiles = ((N + kWgTN - 1) / kWgTN) * ((K + kWgTK - 1) / kWgTK);
  int S = (2 * num_sms() + tiles - 1) / tiles;
  if (S > chunks) S = chunks;
  if (S < 1) S = 1;
  return S;
}
int wgrad_splits(long long M, int N, int K) { return std::max(wgrad_mma_splits(M, N, K), wgrad_umma_splits(M, N, K)); }
// `part` needs wgrad_splits(M, N, K) * N * K floats.  Rows [M, M rounded up to 32) of both operands must be zero
// (the padding rows of every engine buffer are).  Returns the number of partial matrices written.
int wgrad_s3(View G, View X, long long M, int terms, float* part, cudaStream_t st) {
  if (wgrad_umma_supported(G, X)) return wgrad_umma(G, X, M, terms, part, st);      // tcgen05 (wgrad_umma.cu)
  const int N = G.C, K = X.C;
  const int chunks = (int)((M + kWgChunk - 1) / kWgChunk);
  int S = wgrad_mma_splits(M, N, K);
  const int per = (chunks + S - 1) / S;
  S = (chunks + per - 1) / per;
  const int smem = kWgStages * kWgStageBytes;
  optin_max_smem(k_wgrad_s3<3>);
  optin_max_smem(k_wgrad_s3<1>);
  dim3 grid((N + kWgTN - 1) / kWgTN, (K + kWgTK - 1) / kWgTK, S);
  if (terms == 1) launch(k_wgrad_s3<1>, grid, 256, smem, st, G, X, N, K, chunks, per, part);
  else launch(k_wgrad_s3<3>, grid, 256, smem, st, G, X, N, K, chunks, per, part);
  return S;
}


// ------------------------------------------------------------------ gradient scale, boundary conversions
// scale2[0] = 2^floor(peak_log2 - log2(max |g|)) (1 when g is all zero or not finite), scale2[1] = 1 / scale2[0]:
// the incoming gradient enters the fp16 split planes multiplied by scale2[0], every result leaves multiplied by scale2[1].
__global__ void __launch_bounds__(256) k_absmax_part(const float* __restrict__ g, long long n4, float* __restrict__ part) {
  pdl_prologue_done();
  float m = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(g)[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
  }
  __shared__ float red[256];
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void k_grad_scale(const float* __restrict__ part, int n, const float* __restrict__ tail, int ntail, float peak_log2,
                             float* __restrict__ scale2) {
  pdl_prologue_done();
  __shared__ float red[256];
  float m = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, part[i]);
  for (int i = threadIdx.x; i < ntail; i += 256) m = fmaxf(m, fabsf(tail[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float a = red[0];
    float s = 1.0f;
    if (a > 0.0f && a < 3.0e38f) s = exp2f(fminf(fmaxf(floorf(peak_log2 - log2f(a)), -100.0f), 100.0f));
    scale2[0] = s;
    scale2[1] = 1.0f / s;
  }
}
void grad_scale(const float* g, long long n, float peak_log2, float* part, float* scale2, cudaStream_t st) {
  const long long n4 = n / 4;
  int blocks = (int)std::min<long long>((n4 + 255) / 256, 8LL * num_sms());
  if (blocks < 1) blocks = 1;
  launch(k_absmax_part, blocks, 256, 0, st, g, n4, part);
  launch(k_grad_scale, 1, 256, 0, st, (const float*)part, blocks, g + n4 * 4, (int)(n - n4 * 4), peak_log2, scale2);
}

// NCHW fp32 -> split planes, times chan[c] (or 1) times *gscale (or 1).  Thread = (pixel, 8 channels), pixel fastest.
__global__ void k_nchw_to_s3_scaled(const float* __restrict__ x, View out, int C, long long HW, const float* __restrict__ chan,
                                    const float* __restrict__ gscale) {
  pdl_prologue_done();
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c8 = blockIdx.y, b = blockIdx.z;
  const float gs = gscale ? *gscale : 1.0f;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c8 * 8 + i;
    v[i] = (c < C) ? x[((long long)b * C + c) * HW + p] * (chan ? chan[c] * gs : gs) : 0.0f;
  }
  st3x8(out, (long long)b * HW + p, c8 * 8, v);
}
void nchw_to_s3_scaled(const float* x, View out, int B, int C, int H, int W, const float* chan, const float* gscale,
                       cudaStream_t st) {
  const long long HW = (long long)H * W;
  dim3 grid(cdiv_u(HW, 256), (C + 7) / 8, B);
  launch(k_nchw_to_s3_scaled, grid, 256, 0, st, x, out, C, HW, chan, gscale);
}
__global__ void k_s3_to_nchw_scaled(View in, float* __restrict__ x, int C, long long HW, const float* __restrict__ gscale) {
  pdl_prologue_done();
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c8 = blockIdx.y, b = blockIdx.z;
  const float gs = gscale ? *gscale : 1.0f;
  float v[8];
  ld3x8(in, (long long)b * HW + p, c8 * 8, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c8 * 8 + i;
    if (c < C) x[((long long)b * C + c) * HW + p] = v[i] * gs;
  }
}
void s3_to_nchw_scaled(View in, float* x, int B, int C, int H, int W, const float* gscale, cudaStream_t st) {
  const long long HW = (long long)H * W;
  dim3 grid(cdiv_u(HW, 256), (C + 7) / 8, B);
  launch(k_s3_to_nchw_scaled, grid, 256, 0, st, in, x, C, HW, gscale);
}

// part[chunk][c] = sum over the chunk's pixels (all batch items) of a[b,c,p] * b_[b,c,p]   (NCHW fp32 both)
__global__ void __launch_bounds__(256) k_nchw_dot(const float* __restrict__ a, const float* __restrict__ b_, int B, int C,
                                                  long long HW, long long per, float* __restrict__ part) {
  pdl_prologue_done();
  const int c = blockIdx.x;
  const long long p0 = (long long)blockIdx.y * per;
  long long p1 = p0 + per;
  if (p1 > HW) p1 = HW;
  float s = 0.0f;
  for (int b = 0; b < B; ++b) {
    const float* pa = a + ((long long)b * C + c) * HW;
    const float* pb = b_ + ((long long)b * C + c) * HW;
    for (long long p = p0 + threadIdx.x; p < p1; p += 256) s = fmaf(pa[p], pb[p], s);
  }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[(long long)blockIdx.y * C + c] = red[0];
}
// returns the number of partial rows (C floats each)
int nchw_dot(const float* a, const float* b, int B, int C, long long HW, float* part, int max_parts, cudaStream_t st) {
  int chunks = (int)std::min<long long>((HW + 2047) / 2048, (8LL * num_sms() + C - 1) / C);
  if (chunks > max_parts) chunks = max_parts;
  if (chunks < 1) chunks = 1;
  const long long per = (HW + chunks - 1) / chunks;
  chunks = (int)((HW + per - 1) / per);
  launch(k_nchw_dot, dim3(C, chunks), 256, 0, st, a, b, B, C, HW, per, part);
  return chunks;
}
// out[c] = sum_s part[s][c] / div[c]  (0 where div is 0)
__global__ void k_reduce_div(const float* __restrict__ part, int S, int C, const float* __restrict__ div, float* __restrict__ out) {
  pdl_prologue_done();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.0f;
  for (int k = 0; k < S; ++k) s += part[(long long)k * C + c];
  out[c] = div[c] != 0.0f ? s / div[c] : 0.0f;
}
void reduce_div(const float* part, int S, int C, const float* div, float* out, cudaStream_t st) {
  launch(k_reduce_div, cdiv_u(C, 256), 256, 0, st, part, S, C, div, out);
}

// ------------------------------------------------------------------ k x k convolutions (through the im2col view)
// wt[(tap * cin + ci)][n] = w[n][ci][kh][kw], tap = kh * k + kw: the weight of the data gradient as a (k*k*cin, cout, 1, 1)
// convolution over the gradient rows, in the column order of im2col
__global__ void k_wt_kxk(const float* __restrict__ w, float* __restrict__ wt, int cout, int cin, int k) {
  pdl_prologue_done();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long K = (long long)k * k * cin;
  if (i >= K * cout) return;
  const int n = (int)(i % cout);
  const long long j = i / cout;
  const int ci = (int)(j % cin), tap = (int)(j / cin);
  wt[i] = w[(((long long)n * cin + ci) * k + tap / k) * k + tap % k];
}
void wt_kxk(const float* w, float* wt, int cout, int cin, int k, cudaStream_t st) {
  launch(k_wt_kxk, cdiv_u((long long)k * k * cin * cout, 256), 256, 0, st, w, wt, cout, cin, k);
}
// partial sums [S][cout][k*k*cin] (im2col column order) -> weight gradient (cout, cin, k, k)
__global__ void k_reduce_wgrad_kxk(const float* __restrict__ part, long long stride, int S, float* __restrict__ out, int cout,
                                   int cin, int k, const float* __restrict__ scale_dev) {
  pdl_prologue_done();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long K = (long long)k * k * cin;
  if (i >= K * cout) return;
  float s0 = 0.0f, s1 = 0.0f;
  int q = 0;
  for (; q + 1 < S; q += 2) {
    s0 += part[(long long)q * stride + i];
    s1 += part[(long long)(q + 1) * stride + i];
  }
  if (q < S) s0 += part[(long long)q * stride + i];
  const int n = (int)(i / K);
  const long long j = i % K;
  const int ci = (int)(j % cin), tap = (int)(j / cin);
  out[(((long long)n * cin + ci) * k + tap / k) * k + tap % k] = (s0 + s1) * (scale_dev ? *scale_dev : 1.0f);
}
void reduce_wgrad_kxk(const float* part, long long stride, int S, float* out, int cout, int cin, int k, const float* scale_dev,
                      cudaStream_t st) {
  launch(k_reduce_wgrad_kxk, cdiv_u((long long)k * k * cin * cout, 256), 256, 0, st, part, stride, S, out, cout, cin, k, scale_dev);
}
// Gradient of im2col: gx[b, c, hi, wi] = scale * sum over the taps that read this pixel of gcol[(b, ho, wo), tap * C + c].
// gcol: split planes [B*Ho*Wo, k*k*C]; gx: NCHW fp32.  Thread = (pixel, 8 channels), pixel fastest (coalesced NCHW side).
__global__ void k_col2im_nchw(View gcol, float* __restrict__ gx, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                              const float* __restrict__ gscale) {
  pdl_prologue_done();
  const long long HW = (long long)H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c8 = blockIdx.y, b = blockIdx.z;
  const int hi = (int)(p / W), wi = (int)(p % W);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int kh = 0; kh < k; ++kh) {
    const int th = hi + pad - kh;
    if (th < 0 || th % stride) continue;
    const int ho = th / stride;
    if (ho >= Ho) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int tw = wi + pad - kw;
      if (tw < 0 || tw % stride) continue;
      const int wo = tw / stride;
      if (wo >= Wo) continue;
      float v[8];
      ld3x8(gcol, ((long long)b * Ho + ho) * Wo + wo, (kh * k + kw) * C + c8 * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  const float gs = gscale ? *gscale : 1.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c8 * 8 + i;
    if (c < C) gx[((long long)b * C + c) * HW + p] = acc[i] * gs;
  }
}
void col2im_nchw(View gcol, float* gx, int B, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo, const float* gscale,
                 cudaStream_t st) {
  dim3 grid(cdiv_u((long long)H * W, 256), (C + 7) / 8, B);
  launch(k_col2im_nchw, grid, 256, 0, st, gcol, gx, C, H, W, k, stride, pad, Ho, Wo, gscale);
}

// ------------------------------------------------------------------ quantisation / likelihood, training mode
// inference.py:16-27.  mode 0 "ste": out = round(x) (the gradient passes through unchanged: nothing to compute);
// mode 1 "noise": out = x + noise, noise ~ U(-half_bin, half_bin) drawn by the caller (torch's generator, so that a
// seeded run reproduces the reference's).
__global__ void k_quant_train(const float* __restrict__ x, const float* __restrict__ noise, float* __restrict__ out, long long n,
                              int mode) {
  pdl_prologue_done();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = mode == 0 ? rintf(x[i]) : add_rn(x[i], noise[i]);
}
void quant_train(const float* x, const float* noise, float* out, long long n, int mode, cudaStream_t st) {
  launch(k_quant_train, cdiv_u(n, 256), 256, 0, st, x, noise, out, n, mode);
}

// Gradient of the Gaussian likelihood bits (forward values: k_gaussian_bits in kernels.cu).
//   formula 0  models/common_model.py:30-42     sg = clamp(sigma, 1e-5, 1e10); p = Phi((s+.5)/sg) - Phi((s-.5)/sg);
//                                                bits = max(-log2(p + 1e-5), 0)
//   formula 1  refactor/common_model.py:37-68   s clamped to +-6 (seg_video_model.py:347), z clamped to +-12,
//                                                bits = -log2(max(p, 1e-9))
// gs = go * dbits/ds, gsig = go * dbits/dsigma; every clamp passes the gradient on its closed interval and blocks it
// outside, as torch's clamp does.  Evaluated in fp64 (the differences of cdf values cancel).
__global__ void k_gaussian_bits_bwd(const float* __restrict__ sym, const float* __restrict__ sigma, const float* __restrict__ go,
                                    float* __restrict__ gsym, float* __restrict__ gsig, long long n, int formula) {
  pdl_prologue_done();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double kInvSqrt2 = 0.70710678118654752440, kInvSqrt2Pi = 0.39894228040143267794, kLn2 = 0.69314718055994530942;
  double s = sym[i], sg = sigma[i];
  bool s_live = true, sg_live = (sg >= 1e-5 && sg <= 1e10);
  if (formula) {
    if (!(fabs(s) <= 6.0)) s_live = false;
    s = fmin(fmax(s, -6.0), 6.0);
  }
  sg = fmin(fmax(sg, 1e-5), 1e10);
  double zh = (s + 0.5) / sg, zl = (s - 0.5) / sg;
  bool zh_live = true, zl_live = true;
  if (formula) {
    zh_live = fabs(zh) <= 12.0;
    zl_live = fabs(zl) <= 12.0;
    zh = fmin(fmax(zh, -12.0), 12.0);
    zl = fmin(fmax(zl, -12.0), 12.0);
  }
  const double p = 0.5 * (erf(zh * kInvSqrt2) - erf(zl * kInvSqrt2));
  double dbdp;      // d bits / d p
  if (formula) dbdp = (p >= 1e-9) ? -1.0 / (p * kLn2) : 0.0;
  else dbdp = (-log(p + 1e-5) / kLn2 >= 0.0) ? -1.0 / ((p + 1e-5) * kLn2) : 0.0;
  const double ph = zh_live ? kInvSqrt2Pi * exp(-0.5 * zh * zh) : 0.0;     // dp/dzh
  const double pl = zl_live ? kInvSqrt2Pi * exp(-0.5 * zl * zl) : 0.0;     // -dp/dzl
  const double dpds = (ph - pl) / sg;
  const double dpdsg = (-(zh * ph) + zl * pl) / sg;
  const double g = go[i];
  gsym[i] = (float)(s_live ? g * dbdp * dpds : 0.0);
  gsig[i] = (float)(sg_live ? g * dbdp * dpdsg : 0.0);
}
void gaussian_bits_bwd(const float* sym, const float* sigma, const float* go, float* gsym, float* gsig, long long n,
                       int formula, cudaStream_t st) {
  launch(k_gaussian_bits_bwd, cdiv_u(n, 256), 256, 0, st, sym, sigma, go, gsym, gsig, n, formula);
}

}  // namespace dmc
