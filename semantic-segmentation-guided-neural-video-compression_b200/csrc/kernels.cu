// Element-wise / stencil / entropy-model kernels of the DMC engine (sm_100a).
// All of them are HBM-bound: one thread owns 8 consecutive channels (16 B per S3 plane) so
// every global access is a 128-bit transaction, rows are contiguous in the channel dimension
// and consecutive threads touch consecutive 16 B chunks.
#include "kernels.h"

#include <math.h>
#include <stdlib.h>

namespace dmc {

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

static long long g_launches = 0;
void note_launch() { ++g_launches; }
// DMC_PDL=0 / 1 forces programmatic dependent launch off / on; otherwise the engine switches it per frame size
// (pdl_set_auto): on small frames the ~100 launches of a forward are bound by their fixed cost and overlapping the
// next kernel's prologue with the previous one's tail is worth 11 % (128x192: 1.61 -> 1.45 ms per P frame); at
// 1920x1280 it is neutral (9 225 vs 9 229 k clocks per frame).
static int g_pdl_env = -2, g_pdl_auto = 0;
bool pdl_enabled() {
  if (g_pdl_env == -2) {
    const char* e = getenv("DMC_PDL");
    g_pdl_env = !e ? -1 : (e[0] == '1' ? 1 : 0);
  }
  return g_pdl_env >= 0 ? g_pdl_env == 1 : g_pdl_auto == 1;
}
void pdl_set_auto(bool on) { g_pdl_auto = on ? 1 : 0; }
long long launch_count() { return g_launches; }
void add_launches(long long n) { g_launches += n; }

int num_sms() {
  static int n[64] = {};          // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// ------------------------------------------------------------------ accumulate-truncation compensation
// tcgen05.mma adds the 16 products of a k step and the fp32 accumulator at a common alignment and TRUNCATES the sum
// toward zero when it writes the accumulator back (fp32 FMA rounds to nearest).  Measured on B200 against fp64
// (tests/diag/acc_bias.py, profiles/tcgen05_accumulate_bias_r01.txt), zero-mean operands: relative bias of the result
//   K = 64: -1.45   256: -4.69   512: -9.05   1024: -17.67   units of 2^-24   ==  -0.276 * (K/16 + 1)
// i.e. half an accumulator ulp per MMA step, weighted by how far the running sum has grown -- linear in the number
// of accumulate steps, independent of the data (the fp32-FMA kernel: 0.00).  The noise of the two paths is the same
// (rms 7 units at K = 256); what the truncation adds is a MEAN shrink that is coherent across the ~100 layers of a
// frame, and that is what flipped quantised symbols (14 of 2 457 600 on the full-size intra frame against 0 for
// fp32 FMA).  The epilogues therefore scale the main accumulator back by (1 + kappa * (K/16 + 1) * 2^-24), folded
// into the join of the two accumulators (one extra FMA per element).  DMC_ACC_COMP=<kappa> overrides (0 = off).
static float g_acc_kappa = -1.0f;
float acc_comp_kappa() {
  if (g_acc_kappa < 0.0f) {
    const char* e = getenv("DMC_ACC_COMP");
    g_acc_kappa = e ? (float)atof(e) : 0.276f;
    if (!(g_acc_kappa >= 0.0f)) g_acc_kappa = 0.0f;
  }
  return g_acc_kappa;
}
void acc_comp_set_kappa(float k) { g_acc_kappa = k >= 0.0f ? k : 0.0f; }
float acc_comp_scaled(int K) {
  const float steps = (float)((K + 15) / 16 + 1);
  return acc_comp_kappa() * steps * (kLoScale / 16777216.0f);      // kappa * steps * 2^-24, pre-scaled by 2^11
}

// ------------------------------------------------------------------ weights
// `out`: row-major split planes [2][Npad][Kld] (SIMT kernel).  `outb`: the same values tile-blocked for the tcgen05
// kernels, [2][Kld/16][Npad][16] with the 16-byte halves of a row swapped where bit 2 of the row is set -- exactly
// the shared-memory image of a K-major SWIZZLE_32B operand, so a W tile is fetched as a few 512-byte segments.
__global__ void k_pack_gemm_weight(const float* __restrict__ w, h16* __restrict__ out, h16* __restrict__ outb,
                                   int cout, int cin, int kh, int kw, int Npad, int Kld, int K, int pack,
                                   int Cg, int Cg_pad, int transposed) {
  pdl_prologue_done();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Npad * Kld) return;
  int r = (int)(idx / Kld), kk = (int)(idx % Kld);
  int n = -1;
  if (pack == PACK_PAIR) {
    int c2 = cout / 2;
    int ch = (r >> 6) * 32 + (r & 31);
    if (ch < c2) n = ((r & 63) < 32) ? ch : c2 + ch;
  } else if (pack == PACK_SHUF2) {
    int g = r / Cg_pad, c = r % Cg_pad;
    if (c < Cg && g < 4) n = c * 4 + g;
  } else if (r < cout) {
    n = r;
  }
  float v = 0.0f;
  if (n >= 0 && kk < K) {
    int ci = kk % cin, tap = kk / cin;
    int y = tap / kw, x = tap % kw;
    // transposed (1x1 only): `w` is the (cin, cout, 1, 1) weight of the forward convolution, read as its transpose --
    // the weight of the convolution's data gradient (train mode)
    v = transposed ? w[(long long)ci * cout + n] : w[(((long long)n * cin + ci) * kh + y) * kw + x];
  }
  h16 h, l;
  split2(v, h, l);
  long long ps = (long long)Npad * Kld;
  out[idx] = h;
  out[ps + idx] = l;
  if (outb) {
    const int blk = kk >> 4, half = (kk >> 3) & 1, e8 = kk & 7;
    const long long ib = ((long long)blk * Npad + r) * 16 + ((half ^ ((r >> 2) & 1)) << 3) + e8;
    outb[ib] = h;
    outb[ps + ib] = l;
  }
}

__global__ void k_pack_bias(const float* __restrict__ b, float* __restrict__ out, int cout, int Npad,
                            int pack, int Cg, int Cg_pad) {
  pdl_prologue_done();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Npad) return;
  int n = -1;
  if (pack == PACK_PAIR) {
    int c2 = cout / 2;
    int ch = (r >> 6) * 32 + (r & 31);
    if (ch < c2) n = ((r & 63) < 32) ? ch : c2 + ch;
  } else if (pack == PACK_SHUF2) {
    int g = r / Cg_pad, c = r % Cg_pad;
    if (c < Cg && g < 4) n = c * 4 + g;
  } else if (r < cout) {
    n = r;
  }
  out[r] = (n >= 0 && b) ? b[n] : 0.0f;
}

void pack_gemm_weight(const float* w, int cout, int cin, int kh, int kw, const GemmW& g,
                      cudaStream_t st, bool transposed) {
  long long n = (long long)g.Npad * g.Kld;
  launch(k_pack_gemm_weight, cdiv(n, 256), 256, 0, st, w, g.w, g.wb, cout, cin, kh, kw, g.Npad, g.Kld,
                                                   g.K, g.pack, g.Cg, g.Cg_pad, transposed ? 1 : 0);
}
void pack_gemm_bias(const float* bias, int cout, const GemmW& g, cudaStream_t st) {
  launch(k_pack_bias, cdiv(g.Npad, 256), 256, 0, st, bias, g.bias, cout, g.Npad, g.pack, g.Cg, g.Cg_pad);
}

__global__ void k_pack_dw(const float* __restrict__ w, float* __restrict__ out, int C) {
  pdl_prologue_done();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  int tap = i / C, c = i % C;
  out[i] = w[c * 9 + tap];
}
void pack_dw_weight(const float* w, float* out9c, int C, cudaStream_t st) {
  launch(k_pack_dw, cdiv(9 * C, 256), 256, 0, st, w, out9c, C);
}

// ------------------------------------------------------------------ layout conversion
// The kernels that touch a caller-owned tensor take the pointer directly or, when `slot` is given, read it from a
// device-resident slot at run time: a frame captured once as a CUDA graph then runs on whatever tensors the caller
// passes to the next forward (engine.cu writes the slots before every launch).
template <class T>
__device__ __forceinline__ T* io_ptr(T* p, const void* const* slot) {
  return slot ? reinterpret_cast<T*>(const_cast<void*>(*slot)) : p;
}
__global__ void k_unshuffle8_in(const float* x, View out, int B, int Cimg, int H, int W, const void* const* slot) {
  pdl_prologue_done();
  x = io_ptr(x, slot);
  int W8 = W / 8, H8 = H / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * H8 * Cimg * 8 * W8;
  if (idx >= total) return;
  int w8 = (int)(idx % W8);
  long long t = idx / W8;
  int dy = (int)(t % 8); t /= 8;
  int c = (int)(t % Cimg); t /= Cimg;
  int h8 = (int)(t % H8);
  int b = (int)(t / H8);
  const float* src = x + (((long long)b * Cimg + c) * H + (h8 * 8 + dy)) * W + w8 * 8;
  float4 a = *reinterpret_cast<const float4*>(src);
  float4 d = *reinterpret_cast<const float4*>(src + 4);
  float v[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
  long long m = ((long long)b * H8 + h8) * W8 + w8;
  st3x8(out, m, c * 64 + dy * 8, v);
}
void unshuffle8_in(const float* x, View out, int B, int Cimg, int H, int W, cudaStream_t st, const void* const* slot) {
  long long total = (long long)B * (H / 8) * Cimg * 8 * (W / 8);
  launch(k_unshuffle8_in, cdiv(total, 256), 256, 0, st, x, out, B, Cimg, H, W, slot);
}

__global__ void k_shuffle8_out(const float* __restrict__ in, int ld, float* x, int B,
                               int Cimg, int H, int W, const void* const* slot) {
  pdl_prologue_done();
  x = io_ptr(x, slot);
  int W8 = W / 8, H8 = H / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Cimg * H * W8;
  if (idx >= total) return;
  int w8 = (int)(idx % W8);
  long long t = idx / W8;
  int h = (int)(t % H); t /= H;
  int c = (int)(t % Cimg);
  int b = (int)(t / Cimg);
  int h8 = h / 8, dy = h % 8;
  long long m = ((long long)b * H8 + h8) * W8 + w8;
  const float* src = in + m * ld + c * 64 + dy * 8;
  float4 a = *reinterpret_cast<const float4*>(src);
  float4 d = *reinterpret_cast<const float4*>(src + 4);
  a.x = fminf(fmaxf(a.x, 0.f), 1.f); a.y = fminf(fmaxf(a.y, 0.f), 1.f);
  a.z = fminf(fmaxf(a.z, 0.f), 1.f); a.w = fminf(fmaxf(a.w, 0.f), 1.f);
  d.x = fminf(fmaxf(d.x, 0.f), 1.f); d.y = fminf(fmaxf(d.y, 0.f), 1.f);
  d.z = fminf(fmaxf(d.z, 0.f), 1.f); d.w = fminf(fmaxf(d.w, 0.f), 1.f);
  float* dst = x + (((long long)b * Cimg + c) * H + h) * W + w8 * 8;
  *reinterpret_cast<float4*>(dst) = a;
  *reinterpret_cast<float4*>(dst + 4) = d;
}
void shuffle8_out(const float* in, int ld, float* x, int B, int Cimg, int H, int W, cudaStream_t st, const void* const* slot) {
  long long total = (long long)B * Cimg * H * (W / 8);
  launch(k_shuffle8_out, cdiv(total, 256), 256, 0, st, in, ld, x, B, Cimg, H, W, slot);
}

// NCHW fp32 <-> S3 rows.  Thread = (pixel, 8 channels), pixel fastest so the NCHW side is coalesced.
__global__ void k_nchw_to_s3(const float* x, View out, int C, long long HW, const void* const* slot) {
  pdl_prologue_done();
  x = io_ptr(x, slot);
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int c8 = blockIdx.y, b = blockIdx.z;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = c8 * 8 + i;
    v[i] = (c < C) ? x[((long long)b * C + c) * HW + p] : 0.0f;
  }
  st3x8(out, (long long)b * HW + p, c8 * 8, v);
}
void nchw_to_s3(const float* x, View out, int B, int C, int H, int W, cudaStream_t st, const void* const* slot) {
  long long HW = (long long)H * W;
  dim3 grid(cdiv(HW, 256), (C + 7) / 8, B);
  launch(k_nchw_to_s3, grid, 256, 0, st, x, out, C, HW, slot);
}
__global__ void k_s3_to_nchw(View in, float* x, int C, long long HW, const void* const* slot) {
  pdl_prologue_done();
  x = io_ptr(x, slot);
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int c8 = blockIdx.y, b = blockIdx.z;
  float v[8];
  ld3x8(in, (long long)b * HW + p, c8 * 8, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = c8 * 8 + i;
    if (c < C) x[((long long)b * C + c) * HW + p] = v[i];
  }
}
void s3_to_nchw(View in, float* x, int B, int C, int H, int W, cudaStream_t st, const void* const* slot) {
  long long HW = (long long)H * W;
  dim3 grid(cdiv(HW, 256), (C + 7) / 8, B);
  launch(k_s3_to_nchw, grid, 256, 0, st, in, x, C, HW, slot);
}
__global__ void k_f32rows_to_nchw(const float* __restrict__ in, int ld, float* __restrict__ x, int C,
                                  long long HW) {
  pdl_prologue_done();
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int c = blockIdx.y, b = blockIdx.z;
  x[((long long)b * C + c) * HW + p] = in[((long long)b * HW + p) * ld + c];
}
void f32rows_to_nchw(const float* in, int ld, float* x, int B, int C, int H, int W, cudaStream_t st) {
  long long HW = (long long)H * W;
  dim3 grid(cdiv(HW, 256), C, B);
  launch(k_f32rows_to_nchw, grid, 256, 0, st, in, ld, x, C, HW);
}

__global__ void k_scale_cols(View in, const float* __restrict__ scale, View out, long long M, int C8) {
  pdl_prologue_done();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * C8) return;
  long long m = idx / C8;
  int c = (int)(idx % C8) * 8;
  float v[8];
  ld3x8(in, m, c, v);
  if (scale) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = mul_rn(v[i], scale[c + i]);
  }
  st3x8(out, m, c, v);
}
void scale_cols(View in, const float* scale, View out, long long M, cudaStream_t st) {
  int C8 = (in.C + 7) / 8;   // a ragged tail stays inside the row pitch (ld is a multiple of 8)
  launch(k_scale_cols, cdiv(M * C8, 256), 256, 0, st, in, scale, out, M, C8);
}
void copy_view(View in, View out, long long M, cudaStream_t st) { scale_cols(in, nullptr, out, M, st); }

// out[b, h, w, :] = in[b, min(h, Hin-1), min(w, Win-1), :]: replicate padding on the right / bottom when the output
// grid is larger (inference.py:40-43), a crop to the top-left corner when it is smaller (`[:, :, :h, :w]`).
__global__ void k_regrid(View in, int Hin, int Win, View out, int Hout, int Wout, int B, int C8) {
  pdl_prologue_done();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Hout * Wout * C8) return;
  int c = (int)(idx % C8) * 8;
  long long t = idx / C8;
  int w = (int)(t % Wout);
  t /= Wout;
  int h = (int)(t % Hout);
  long long b = t / Hout;
  const long long src = (b * Hin + min(h, Hin - 1)) * Win + min(w, Win - 1);
  const h16* q = in.p + s3_unit_offset(in, src, c);
  h16* d = out.p + s3_unit_offset(out, (b * Hout + h) * Wout + w, c);
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(q);
  *reinterpret_cast<uint4*>(d + out.ps) = *reinterpret_cast<const uint4*>(q + in.ps);
}
void regrid(View in, int Hin, int Win, View out, int Hout, int Wout, int B, cudaStream_t st) {
  int C8 = (in.C + 7) / 8;   // a ragged tail stays inside the last 16-column block
  long long n = (long long)B * Hout * Wout * C8;
  launch(k_regrid, cdiv(n, 256), 256, 0, st, in, Hin, Win, out, Hout, Wout, B, C8);
}

// One launch checks up to 8 tensors (blockIdx.y = tensor): bit i of *flag is set if tensor i holds a non-finite value
// or one saturated at the fp16 range limit of the split storage format (|x| >= 65504 cannot be represented).
__global__ void k_finite_check(FiniteList l, int* flag) {
  pdl_prologue_done();
  const int t = blockIdx.y;
  const View v = l.v[t];
  const int C8 = v.C / 8;
  const long long n = l.M[t] * C8;
  bool bad = false;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / C8;
    const int c = (int)(idx % C8) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(v.p + s3_unit_offset(v, m, c));
    const uint32_t* u = &a.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) bad |= ((u[i] & 0x7fffu) >= 0x7bffu) | ((u[i] & 0x7fff0000u) >= 0x7bff0000u);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1 << t);
}
void finite_check(const FiniteList& l, int* flag, cudaStream_t st) {
  if (l.n < 1) return;
  long long most = 0;
  for (int i = 0; i < l.n; ++i) most = l.M[i] * (l.v[i].C / 8) > most ? l.M[i] * (l.v[i].C / 8) : most;
  unsigned gx = cdiv(most, 256 * 8);
  if (gx < 1) gx = 1;
  launch(k_finite_check, dim3(gx, l.n), 256, 0, st, l, flag);
}

// copies the engine's flag word to the caller's (if the caller gave one: *slot may be null)
__global__ void k_copy_flag(const int* src, const void* const* slot) {
  pdl_prologue_done();
  int* dst = reinterpret_cast<int*>(const_cast<void*>(*slot));
  if (dst) *dst = *src;
}
void copy_flag(const int* src, const void* const* slot, cudaStream_t st) { launch(k_copy_flag, 1, 1, 0, st, src, slot); }

__global__ void k_set_io(IoSlots* dst, IoSlots v) { *dst = v; }
void set_io(IoSlots* dst, const IoSlots& v, cudaStream_t st) {
  note_launch();
  k_set_io<<<1, 1, 0, st>>>(dst, v);
}

// ------------------------------------------------------------------ depthwise 3x3 (layers.py:56)
// Depthwise 3x3 (layers.py:59-61, groups == C).  One thread produces 8 channels of kDwPix horizontally
// adjacent pixels: each of the three input rows is loaded once as kDwPix + 2 columns (instead of 3 per
// output), which cuts the L1 traffic and the S3 joins 2.4x.  The taps of one output are accumulated in
// the order (dy, dx) = (-1,-1) ... (1,1) with the bias added last; out-of-image taps contribute
// fmaf(0, w, acc) == acc, so the result does not depend on the blocking.
constexpr int kDwPix = 4;
// kF32In: the input is fp32 rows [M, ld] (no S3 join: the producing contraction writes fp32 for this consumer)
// A CTA covers kDwTR image rows x kDwTP pixel groups (x all channels): its (kDwTR + 2) x (4 kDwTP + 2) input pixels
// are fetched from L2 once and the 3 x 3 reuse is served by L1 (one image row per CTA fetched every input row three
// times: 3.4x the unique bytes over L2).
constexpr int kDwTP = 2;
template <bool kF32In>
__global__ void __launch_bounds__(512)
k_dwconv3x3(View in, const float* __restrict__ in32, int ld32, const float* __restrict__ w9c,
            const float* __restrict__ bias, View out, int B, int H, int W, int C8, int WG, int TR, int dbg) {
  pdl_prologue_done();
  const int cg = (int)threadIdx.x % C8;
  const int pos = (int)threadIdx.x / C8;
  const int WT = (WG + kDwTP - 1) / kDwTP, HT = (H + TR - 1) / TR;
  long long t = blockIdx.x;
  const int wt = (int)(t % WT);
  t /= WT;
  const int ht = (int)(t % HT);
  const long long b = t / HT;
  const int wg = wt * kDwTP + pos % kDwTP;
  const int h = ht * TR + pos / kDwTP;
  if (h >= H || wg >= WG || pos >= TR * kDwTP) return;
  const int c = cg * 8, C = C8 * 8;
  const int w0 = wg * kDwPix;
  float acc[kDwPix][8];
#pragma unroll
  for (int p = 0; p < kDwPix; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.0f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int hi = h + dy - 1;
    if ((unsigned)hi >= (unsigned)H) continue;            // a whole tap row outside: acc unchanged
    const long long row = (b * H + hi) * W;
    float v[kDwPix + 2][8];
#pragma unroll
    for (int j = 0; j < kDwPix + 2; ++j) {
      const int wi = w0 + j - 1;
      if ((unsigned)wi < (unsigned)W && !(dbg & 2)) {
        if (kF32In) {
          const float4* q = reinterpret_cast<const float4*>(in32 + (row + wi) * ld32 + c);
          const float4 a = q[0], d = q[1];
          v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
          v[j][4] = d.x; v[j][5] = d.y; v[j][6] = d.z; v[j][7] = d.w;
        } else {
          ld3x8(in, row + wi, c, v[j]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] = 0.0f;
      }
    }
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const float* wt = w9c + (dy * 3 + dx) * C + c;
      const float4 wa = __ldg(reinterpret_cast<const float4*>(wt));
      const float4 wb = __ldg(reinterpret_cast<const float4*>(wt + 4));
#pragma unroll
      for (int p = 0; p < kDwPix; ++p) {
        acc[p][0] = fmaf(v[p + dx][0], wa.x, acc[p][0]); acc[p][1] = fmaf(v[p + dx][1], wa.y, acc[p][1]);
        acc[p][2] = fmaf(v[p + dx][2], wa.z, acc[p][2]); acc[p][3] = fmaf(v[p + dx][3], wa.w, acc[p][3]);
        acc[p][4] = fmaf(v[p + dx][4], wb.x, acc[p][4]); acc[p][5] = fmaf(v[p + dx][5], wb.y, acc[p][5]);
        acc[p][6] = fmaf(v[p + dx][6], wb.z, acc[p][6]); acc[p][7] = fmaf(v[p + dx][7], wb.w, acc[p][7]);
      }
    }
  }
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int p = 0; p < kDwPix; ++p) {
    if (w0 + p >= W) break;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = add_rn(acc[p][i], bv[i]);
    if (!(dbg & 1) || o[0] == 123.456f) st3x8(out, (b * H + h) * W + w0 + p, c, o);
  }
}
// rows per CTA tile: as many as keep the CTA at <= 512 threads (C8 * kDwTP threads per row), at most 8
static int dw_tile_rows(int C8) {
  static int cap = 0;
  if (!cap) {
    const char* v = getenv("DMC_DW_TR");       // experiments: rows per CTA tile
    cap = v ? atoi(v) : 2;
    if (cap < 1) cap = 1;
  }
  int tr = 512 / (C8 * kDwTP);
  if (tr > cap) tr = cap;
  if (tr < 1) tr = 1;
  return tr;
}
// probe switch of the one-row kernel (DMC_DW_DBG: 1 no stores, 2 no input loads): 28.8 / 22.7 / 18.6 / 14.5 us at
// 160x240x256 for 0 / 1 / 2 / 3 -- half of its time is instruction issue, which is what the strip kernel removes
static int dw_dbg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DMC_DW_DBG"); v = e ? atoi(e) : 0; }
  return v;
}
static unsigned dw_grid(int B, int H, int WG, int TR) {
  return (unsigned)((long long)B * ((H + TR - 1) / TR) * ((WG + kDwTP - 1) / kDwTP));
}
void dwconv3x3(View in, const float* w9c, const float* bias, View out, int B, int H, int W,
               cudaStream_t st) {
  const int C8 = in.C / 8;
  const int WG = (W + kDwPix - 1) / kDwPix;
  const int TR = dw_tile_rows(C8);
  launch(k_dwconv3x3<false>, dw_grid(B, H, WG, TR), C8 * kDwTP * TR, 0, st, in, (const float*)nullptr, 0, w9c, bias, out, B, H, W, C8, WG, TR, dw_dbg());
}
// Strip version for fp32 input rows: a thread owns 4 channels x 4 adjacent pixels and walks `strip` image rows
// downwards with the three input rows it needs in registers (a rolling window: every input row is loaded once per
// strip, not once per output row), the nine taps of its channels loaded once.  The kernel above spends 34
// instructions per output -- 18 tap loads and 36 input loads per 32 outputs, predicates, zero fills -- and half of
// its time is there (14.5 of 28.8 us at 160x240x256 with loads and stores switched off); this one spends ~16.
// Same order of the nine FMAs per output as above, so the results are bit-identical.
#ifndef DMC_DW_MINB
#define DMC_DW_MINB 3
#endif
struct DwRow { float v[kDwPix + 2][4]; };
__device__ __forceinline__ void dw_load_row(DwRow& r, const float* __restrict__ in32, int ld, long long rowbase, int h,
                                            int H, int w0, int W, int c) {
  if ((unsigned)h >= (unsigned)H) {
#pragma unroll
    for (int j = 0; j < kDwPix + 2; ++j) { r.v[j][0] = r.v[j][1] = r.v[j][2] = r.v[j][3] = 0.0f; }
    return;
  }
  const float* q = in32 + ((rowbase + h) * W + w0 - 1) * ld + c;
#pragma unroll
  for (int j = 0; j < kDwPix + 2; ++j) {
    const int wi = w0 + j - 1;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((unsigned)wi < (unsigned)W) a = *reinterpret_cast<const float4*>(q + (long long)j * ld);
    r.v[j][0] = a.x; r.v[j][1] = a.y; r.v[j][2] = a.z; r.v[j][3] = a.w;
  }
}
__global__ void __launch_bounds__(128, DMC_DW_MINB)
k_dwconv3x3_strip(const float* __restrict__ in32, int ld, const float* __restrict__ w9c,
                  const float* __restrict__ bias, View out, int B, int H, int W, int C4, int WG, int HS, int strip) {
  pdl_prologue_done();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HS * WG * C4) return;
  const int c4 = (int)(idx % C4);
  long long t = idx / C4;
  const int wg = (int)(t % WG);
  t /= WG;
  const int hs = (int)(t % HS);
  const long long b = t / HS;
  const int c = c4 * 4, C = C4 * 4, w0 = wg * kDwPix;
  const int h0 = hs * strip, h1 = min(h0 + strip, H);
  float wt[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w9c + k * C + c));
    wt[k][0] = a.x; wt[k][1] = a.y; wt[k][2] = a.z; wt[k][3] = a.w;
  }
  const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
  const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
  const long long rowbase = b * H;
  // output row h from rows (top, mid, bot) = (h-1, h, h+1)
  auto emit = [&](const DwRow& top, const DwRow& mid, const DwRow& bot, int h) {
    const DwRow* rows[3] = {&top, &mid, &bot};
#pragma unroll
    for (int p = 0; p < kDwPix; ++p) {
      if (w0 + p >= W) break;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(rows[dy]->v[p + dx][i], wt[dy * 3 + dx][i], acc[i]);
      uint32_t hi0, lo0, hi1, lo1;
      split2x2(add_rn(acc[0], bv[0]), add_rn(acc[1], bv[1]), hi0, lo0);
      split2x2(add_rn(acc[2], bv[2]), add_rn(acc[3], bv[3]), hi1, lo1);
      h16* q = out.p + s3_unit_offset(out, (rowbase + h) * W + w0 + p, c) + (c & 4);
      *reinterpret_cast<uint2*>(q) = make_uint2(hi0, hi1);
      *reinterpret_cast<uint2*>(q + out.ps) = make_uint2(lo0, lo1);
    }
  };
  DwRow r0, r1, r2;
  dw_load_row(r0, in32, ld, rowbase, h0 - 1, H, w0, W, c);
  dw_load_row(r1, in32, ld, rowbase, h0, H, w0, W, c);
  for (int h = h0; h < h1; h += 3) {
    dw_load_row(r2, in32, ld, rowbase, h + 1, H, w0, W, c);
    emit(r0, r1, r2, h);
    if (h + 1 >= h1) break;
    dw_load_row(r0, in32, ld, rowbase, h + 2, H, w0, W, c);
    emit(r1, r2, r0, h + 1);
    if (h + 2 >= h1) break;
    dw_load_row(r1, in32, ld, rowbase, h + 3, H, w0, W, c);
    emit(r2, r0, r1, h + 2);
  }
}

// (Tried and dropped: the same taps from a shared-memory tile filled by cp.async.bulk -- (4 + 2) row segments of
// 10 pixels x C floats on one mbarrier, 2 CTAs of 256 threads per SM: 35 us against 29 us at 160x240x256.  ncu on
// the register version: issue slots 40 % busy, L1 64 %, DRAM 2.6 TB/s -- it is bound by L1 transactions and
// instruction count, not by latency.  CTA tiles of 2-8 image rows x 8 pixels instead of one row change nothing.
// A persistent version of the shared-memory kernel -- one CTA per SM, 16 warps, two 108 KB tile buffers filled one
// tile ahead, taps in shared memory -- is no better either: 32.9 us, +4 % clocks on the frame.)
// (A mapping with one thread per 16-channel block and consecutive threads along the image row makes the blocked
// stores contiguous but the fp32 reads strided by a whole pixel: 78 us instead of 33 us at 160x240x256.  The
// channel-fastest mapping below keeps every 16-byte store inside a fully written 32-byte sector.)
void dwconv3x3_f32(const float* in, int ld, const float* w9c, const float* bias, View out, int B, int H, int W,
                   cudaStream_t st) {
  static int mode = -1;
  if (mode < 0) {
    const char* v = getenv("DMC_DW_STRIP");       // DMC_DW_STRIP=0: the one-row version (A/B runs)
    mode = (v && v[0] == '0') ? 0 : 1;
  }
  const int WG = (W + kDwPix - 1) / kDwPix;
  if (mode == 1 && ld % 4 == 0) {
    // strips as long as possible while the launch still fills the machine in one wave (157 registers: three
    // 128-thread CTAs per SM), at least 4 rows (a strip reads its rows + 2)
    const int C4 = out.C / 4;
    const long long warps_per_strip = ((long long)B * WG * C4 + 31) / 32;
    long long hs_max = (long long)num_sms() * (4 * DMC_DW_MINB) / warps_per_strip;
    if (hs_max < 1) hs_max = 1;
    int strip = (int)((H + hs_max - 1) / hs_max);
    if (strip < 4) strip = 4;
    if (strip > 32) strip = 32;
    const int HS = (H + strip - 1) / strip;
    const long long n = (long long)B * HS * WG * C4;
    launch(k_dwconv3x3_strip, cdiv(n, 128), 128, 0, st, in, ld, w9c, bias, out, B, H, W, C4, WG, HS, strip);
    return;
  }
  const int C8 = out.C / 8;
  const int TR = dw_tile_rows(C8);
  launch(k_dwconv3x3<true>, dw_grid(B, H, WG, TR), C8 * kDwTP * TR, 0, st, out, in, ld, w9c, bias, out, B, H, W, C8, WG, TR, dw_dbg());
}

// ------------------------------------------------------------------ im2col (k x k, stride, pad)
__global__ void k_im2col(View in, View out, int B, int H, int W, int k, int stride, int pad, int Ho,
                         int Wo, int tap_stride, int col_off, int C8) {
  pdl_prologue_done();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long Mo = (long long)B * Ho * Wo;
  int taps = k * k;
  if (idx >= Mo * taps * C8) return;
  int c = (int)(idx % C8) * 8;
  long long t = idx / C8;
  int tap = (int)(t % taps);
  long long mo = t / taps;
  int wo = (int)(mo % Wo);
  int ho = (int)((mo / Wo) % Ho);
  int b = (int)(mo / ((long long)Wo * Ho));
  int hi = ho * stride - pad + tap / k, wi = wo * stride - pad + tap % k;
  uint4 z = make_uint4(0, 0, 0, 0), a = z, bb = z;
  if ((unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W) {
    const h16* q = in.p + s3_unit_offset(in, ((long long)b * H + hi) * W + wi, c);
    a = *reinterpret_cast<const uint4*>(q);
    bb = *reinterpret_cast<const uint4*>(q + in.ps);
  }
  h16* d = out.p + s3_unit_offset(out, mo, tap * tap_stride + col_off + c);
  *reinterpret_cast<uint4*>(d) = a;
  *reinterpret_cast<uint4*>(d + out.ps) = bb;
}
// Row-fastest variant: a thread moves one 16-column block of one (output row, tap) -- 32 contiguous bytes per
// plane -- and consecutive threads take consecutive output rows, so a warp writes 1 KB contiguous per plane (the
// blocked layout keeps the rows of a column block together) and, at stride 1, reads as much.
__global__ void k_im2col_rows(View in, View out, int B, int H, int W, int k, int stride, int pad, int Ho,
                              int Wo, int tap_stride, int col_off, int C16, unsigned Mo) {
  pdl_prologue_done();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned taps = (unsigned)(k * k);
  if (idx >= Mo * C16) return;
  const unsigned mo = idx % Mo, blk = idx / Mo, tap = blockIdx.y;
  (void)taps;
  const unsigned wo = mo % (unsigned)Wo, t = mo / (unsigned)Wo;
  const unsigned ho = t % (unsigned)Ho, b = t / (unsigned)Ho;
  const int hi = (int)ho * stride - pad + (int)tap / k, wi = (int)wo * stride - pad + (int)tap % k;
  const int c = (int)blk * 16;
  uint4 z = make_uint4(0, 0, 0, 0), a0 = z, a1 = z, b0 = z, b1 = z;
  if ((unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W) {
    const long long r = ((long long)b * H + hi) * W + wi;
    const h16* q0 = in.p + s3_unit_offset(in, r, c);
    const h16* q1 = in.p + s3_unit_offset(in, r, c + 8);
    a0 = *reinterpret_cast<const uint4*>(q0);
    a1 = *reinterpret_cast<const uint4*>(q1);
    b0 = *reinterpret_cast<const uint4*>(q0 + in.ps);
    b1 = *reinterpret_cast<const uint4*>(q1 + in.ps);
  }
  const int dc = (int)tap * tap_stride + col_off + c;
  h16* d0 = out.p + s3_unit_offset(out, mo, dc);
  h16* d1 = out.p + s3_unit_offset(out, mo, dc + 8);
  *reinterpret_cast<uint4*>(d0) = a0;
  *reinterpret_cast<uint4*>(d1) = a1;
  *reinterpret_cast<uint4*>(d0 + out.ps) = b0;
  *reinterpret_cast<uint4*>(d1 + out.ps) = b1;
}
void im2col(View in, View out, int B, int H, int W, int k, int stride, int pad, int Ho, int Wo,
            int tap_stride, int col_off, cudaStream_t st) {
  const long long Mo = (long long)B * Ho * Wo;
  if (in.C % 16 == 0 && tap_stride % 16 == 0 && col_off % 16 == 0 && Mo * (in.C / 16) < (1LL << 31)) {
    const int C16 = in.C / 16;
    dim3 grid(cdiv(Mo * C16, 256), k * k);
    launch(k_im2col_rows, grid, 256, 0, st, in, out, B, H, W, k, stride, pad, Ho, Wo, tap_stride, col_off, C16,
           (unsigned)Mo);
    return;
  }
  int C8 = in.C / 8;
  long long n = (long long)B * Ho * Wo * k * k * C8;
  launch(k_im2col, cdiv(n, 256), 256, 0, st, in, out, B, H, W, k, stride, pad, Ho, Wo, tap_stride,
         col_off, C8);
}

// ------------------------------------------------------------------ fp32 CUDA-core GEMM
// out[m][n] = sum_k A[m][k] * W[n][k]; 64x64 tile, 256 threads, 4x4 outputs per thread with
// columns tx, tx+16, tx+32, tx+48 so that in PACK_PAIR mode (tile = one 64-column group) a
// thread holds both members of every chunk-add pair.  Validation backend + odd shapes.
__global__ void __launch_bounds__(256) k_gemm_simt(View a, const h16* __restrict__ wp, long long wps,
                                                   int Kld, int K, long long M, Epi e) {
  pdl_prologue_done();
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  long long m0 = (long long)blockIdx.x * 64;
  int n0 = blockIdx.y * 64;
  int lr = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    long long ar = m0 + lr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + lk + i;
      As[lk + i][lr] = (ar < M && k < K) ? ld3(a, ar, k) : 0.0f;
      const h16* q = wp + (long long)(n0 + lr) * Kld + k;
      Bs[lk + i][lr] = (k < Kld) ? join2(q[0], q[wps]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float partner = (e.pack == PACK_PAIR && j < 2) ? acc[i][j + 2] : 0.0f;
      epi_store(e, m, n0 + tx + 16 * j, acc[i][j], partner);
    }
  }
}
void gemm_simt(View a, const GemmW& w, const Epi& e, long long M, cudaStream_t st) {
  dim3 grid(cdiv(M, 64), w.Npad / 64);
  launch(k_gemm_simt, grid, 256, 0, st, a, w.w, (long long)w.Npad * w.Kld, w.Kld, w.K, M, e);
}

// ------------------------------------------------------------------ block reduction helper
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (wid == 0)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // valid in thread 0
}

// ------------------------------------------------------------------ likelihoods
// erf as the reference's CPU path computes it.  torch's CPU erf is MKL VML's vsErf (high-accuracy mode): within one
// ulp, equal to the correctly rounded value on 95 % of the inputs -- and saturated to exactly +-1 from
// |x| >= 3.832507 (0x407547cb) on, where the correctly rounded value stays at 1 - 2^-24 up to 3.9192 (measured over
// every fp32 in [3, 4)).  The likelihoods below cancel down to the last bit of erf for tail symbols: in that band a
// symbol costs 29.9 bits with the saturated erf and 25 with the rounded one, which alone moved bpp by 6.5e-4
// relative (0.1 % of the symbols with random-init weights); CUDA's 2-ulp erff moves it by as much again.  So:
// correctly rounded from double, with the reference's saturation point.
__device__ __forceinline__ float erf_cr(float x) {
  if (fabsf(x) >= __uint_as_float(0x407547cbu)) return copysignf(1.0f, x);
  return (float)erf((double)x);
}

__device__ __forceinline__ float bits_old(float s, float sigma) {
  // models/common_model.py:30-42: Normal(0, clamp(sigma)).cdf difference, log(p + 1e-5)
  float sg = fminf(fmaxf(sigma, 1e-5f), 1e10f);
  float inv = 1.0f / sg;
  const float r2 = 1.41421356237309504880f;
  float hi = mul_rn(0.5f, add_rn(1.0f, erf_cr(mul_rn(add_rn(s, 0.5f), inv) / r2)));
  float lo = mul_rn(0.5f, add_rn(1.0f, erf_cr(mul_rn(sub_rn(s, 0.5f), inv) / r2)));
  float p = sub_rn(hi, lo);
  float b = mul_rn(logf(add_rn(p, 1e-5f)), -1.4426950408889634f);
  return fmaxf(b, 0.0f);
}
__device__ __forceinline__ float nan_to_num(float v, float nanv, float pinf, float ninf) {
  if (isnan(v)) return nanv;
  if (isinf(v)) return v > 0 ? pinf : ninf;
  return v;
}
__device__ __forceinline__ float bits_refactor(float s, float sigma) {
  // refactor/common_model.py:37-68; the +-6 clamp is seg_video_model.py:347
  s = fminf(fmaxf(s, -6.0f), 6.0f);
  s = nan_to_num(s, 0.0f, 1e4f, -1e4f);
  float sg = nan_to_num(sigma, 1e-5f, 1e10f, 1e-5f);
  sg = fminf(fmaxf(sg, 1e-5f), 1e10f);
  float inv = 1.0f / sg;
  float zh = fminf(fmaxf(mul_rn(add_rn(s, 0.5f), inv), -12.0f), 12.0f);
  float zl = fminf(fmaxf(mul_rn(sub_rn(s, 0.5f), inv), -12.0f), 12.0f);
  const float r2 = 1.41421356237309504880f;
  float p = mul_rn(0.5f, sub_rn(erf_cr(zh / r2), erf_cr(zl / r2)));
  p = nan_to_num(p, 0.0f, 0.0f, 0.0f);
  p = fmaxf(p, 1e-9f);
  return -log2f(p);
}

__global__ void k_gaussian_bits(const float* __restrict__ sym, const float* __restrict__ sigma,
                                float* __restrict__ bits, long long n, int formula) {
  pdl_prologue_done();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bits[i] = formula ? bits_refactor(sym[i], sigma[i]) : bits_old(sym[i], sigma[i]);
}
void gaussian_bits(const float* sym, const float* sigma, float* bits, long long n, int formula,
                   cudaStream_t st) {
  launch(k_gaussian_bits, cdiv(n, 256), 256, 0, st, sym, sigma, bits, n, formula);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// checkerboard owner step of element (h, w, c)   (models/common_model.py:101-114,152-169)
__device__ __forceinline__ int prior_owner(int scheme, int h, int w, int c, int C) {
  if (scheme == 2) return (h + w + (c >= C / 2 ? 1 : 0)) & 1;
  const int own[4][4] = {{0, 3, 2, 1}, {3, 0, 1, 2}, {2, 1, 0, 3}, {1, 2, 3, 0}};  // [quarter][pos]
  int pos = (h & 1) * 2 + (w & 1);
  return own[c / (C / 4)][pos];
}

// One checkerboard step of compress_prior_2x / _4x (models/common_model.py:81-90,121-149,188-248).
// A thread owns 8 consecutive channels of one latent position: the checkerboard owner is constant over them (the
// channel halves / quarters start at multiples of 64), every S3 access is a 16-byte load / store per plane and the
// fp32 symbol / sigma rows go out as two float4 each.  (The first version had one thread per element: 57 us for the
// ~44 MB a P frame moves through these kernels = 0.77 TB/s.)
__global__ void k_prior_step(PriorArgs a) {
  pdl_prologue_done();
  const int C8 = a.C >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)a.B * a.H * a.W;
  if (idx >= M * C8) return;
  const long long m = idx / C8;
  const int c = (int)(idx % C8) * 8;
  const int w = (int)(m % a.W), h = (int)((m / a.W) % a.H);
  const int owner = prior_owner(a.scheme, h, w, c, a.C);
  if (owner != a.step) {
    if (a.step == 0 && a.mode != 2) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st3x8(a.yh, m, c, z);
    }
    return;
  }
  float sg[8], mu[8], ys[8];
  if (a.mode == 0) {
    float y[8];
    ld3x8(a.y, m, c, y);
    if (a.scheme == 2) {
      float q[8];
      ld3x8(a.params, m, c, q);
#pragma unroll
      for (int i = 0; i < 8; ++i) ys[i] = mul_rn(y[i], 1.0f / fmaxf(q[i], 0.5f));          // inference.py:29-33
    } else {
      const float qe = add_rn(mul_rn(sigmoidf_(ld3(a.params, m, 0)), 1.5f), 0.5f);         // common_model.py:178-180
#pragma unroll
      for (int i = 0; i < 8; ++i) ys[i] = mul_rn(y[i], qe);
    }
  }
  if (a.step == 0) {
    if (a.scheme == 2) {
      ld3x8(a.params, m, a.C + c, sg);
      ld3x8(a.params, m, 2 * a.C + c, mu);
    } else {                                // [qe, qd | sigma0 | mu0]: the two leading columns break the 8-alignment
#pragma unroll
      for (int i = 0; i < 8; ++i) { sg[i] = ld3(a.params, m, 2 + c + i); mu[i] = ld3(a.params, m, 2 + a.C + c + i); }
    }
  } else {
    ld3x8(a.sp, m, c, sg);
    ld3x8(a.sp, m, a.C + c, mu);
  }
  float4* pg = reinterpret_cast<float4*>(a.sig + m * a.C + c);
  pg[0] = make_float4(sg[0], sg[1], sg[2], sg[3]); pg[1] = make_float4(sg[4], sg[5], sg[6], sg[7]);
  if (a.mode == 2) return;
  float s[8], yh[8];
  if (a.mode == 1) {
    const long long HW = (long long)a.H * a.W;
    const float* src = a.sym_in + ((m / HW) * a.C + c) * HW + m % HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = src[i * HW];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = rintf(sub_rn(ys[i], mu[i]));                   // torch.round = half-to-even
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) yh[i] = add_rn(s[i], mu[i]);
  st3x8(a.yh, m, c, yh);
  float4* ps = reinterpret_cast<float4*>(a.sym + m * a.C + c);
  ps[0] = make_float4(s[0], s[1], s[2], s[3]); ps[1] = make_float4(s[4], s[5], s[6], s[7]);
}
void prior_step(const PriorArgs& a, cudaStream_t st) {
  long long n = (long long)a.B * a.H * a.W * (a.C / 8);
  launch(k_prior_step, cdiv(n, 128), 128, 0, st, a);
}

__global__ void k_prior_finish(PriorArgs a, View y_hat, int formula, double* bits_acc) {
  pdl_prologue_done();
  const int C8 = a.C >> 3;
  const long long per = (long long)a.H * a.W * C8;      // 8-channel units per sample
  const int b = blockIdx.y;
  double local = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
       i += (long long)gridDim.x * blockDim.x) {
    const long long unit = (long long)b * per + i;
    const long long m = unit / C8;
    const int c = (int)(unit % C8) * 8;
    float q[8], yh[8], o[8];
    if (a.scheme == 2) {
      ld3x8(a.params, m, c, q);
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = fmaxf(q[k], 0.5f);
    } else {
      const float qd = add_rn(mul_rn(sigmoidf_(ld3(a.params, m, 1)), 1.5f), 0.5f);
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = qd;
    }
    ld3x8(a.yh, m, c, yh);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = mul_rn(yh[k], q[k]);          // inference.py:35-38
    st3x8(y_hat, m, c, o);
    const float4* ps = reinterpret_cast<const float4*>(a.sym + m * a.C + c);
    const float4* pg = reinterpret_cast<const float4*>(a.sig + m * a.C + c);
    const float4 s0 = ps[0], s1 = ps[1], g0 = pg[0], g1 = pg[1];
    const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) local += (double)(formula ? bits_refactor(s[k], g[k]) : bits_old(s[k], g[k]));
  }
  double tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(&bits_acc[b], tot);
}
void prior_finish(const PriorArgs& a, View y_hat, int formula, double* bits_acc, cudaStream_t st) {
  long long per = (long long)a.H * a.W * (a.C / 8);
  unsigned gx = cdiv(per, 128);
  if (gx < 1) gx = 1;
  dim3 grid(gx, a.B);
  launch(k_prior_finish, grid, 128, 0, st, a, y_hat, formula, bits_acc);
}

// Bitparm chain (entropy_models.py:84-106) -> sigmoid (:139-150)
__device__ __forceinline__ float bitparm_cdf(float v, const BitparmRow& t, int c) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float h = t.p[3 * k][c], b = t.p[3 * k + 1][c];
    float sp = (h > 20.0f) ? h : log1pf(expf(h));          // F.softplus
    v = add_rn(mul_rn(v, sp), b);
    if (k < 3) v = add_rn(v, mul_rn(tanhf(v), tanhf(t.p[3 * k + 2][c])));
  }
  return sigmoidf_(v);
}
__global__ void k_round_z_bits(View z, View z_hat, long long per, int C, BitparmRow t,
                               double* bits_acc) {
  pdl_prologue_done();
  int b = blockIdx.y;
  double local = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
       i += (long long)gridDim.x * blockDim.x) {
    long long idx = (long long)b * per + i;
    long long m = idx / C;
    int c = (int)(idx % C);
    float zr = rintf(ld3(z, m, c));
    st3(z_hat, m, c, zr);
    float p = sub_rn(bitparm_cdf(add_rn(zr, 0.5f), t, c), bitparm_cdf(sub_rn(zr, 0.5f), t, c));
    float bits = fmaxf(mul_rn(logf(add_rn(p, 1e-5f)), -1.4426950408889634f), 0.0f);
    local += (double)bits;
  }
  double tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(&bits_acc[b], tot);
}
void round_z_bits(View z, View z_hat, int B, int HW, int C, BitparmRow t, double* bits_acc,
                  cudaStream_t st) {
  long long per = (long long)HW * C;
  unsigned gx = cdiv(per, 256 * 4);
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  launch(k_round_z_bits, grid, 256, 0, st, z, z_hat, per, C, t, bits_acc);
}

__global__ void k_finalize_bpp(const double* by, const double* bz, float* bpp3, int B, float pixels,
                               const void* const* slot) {
  pdl_prologue_done();
  bpp3 = io_ptr(bpp3, slot);
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float y = (float)by[b] / pixels, z = (float)bz[b] / pixels;
  bpp3[3 * b] = add_rn(y, z);
  bpp3[3 * b + 1] = y;
  bpp3[3 * b + 2] = z;
}
void finalize_bpp(const double* by, const double* bz, float* bpp3, int B, int pixels, cudaStream_t st,
                  const void* const* slot) {
  launch(k_finalize_bpp, cdiv(B, 64), 64, 0, st, by, bz, bpp3, B, (float)pixels, slot);
}

// ------------------------------------------------------------------ mask conditioning
__global__ void k_film(View y, View gb, View out, long long M, int C) {
  pdl_prologue_done();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int C8 = C / 8;
  if (idx >= M * C8) return;
  long long m = idx / C8;
  int c = (int)(idx % C8) * 8;
  float v[8], g[8], b[8];
  ld3x8(y, m, c, v);
  ld3x8(gb, m, c, g);
  ld3x8(gb, m, C + c, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = add_rn(mul_rn(v[i], add_rn(1.0f, g[i])), b[i]);  // seg_video_model.py:327-328
  st3x8(out, m, c, v);
}
void film(View y, View gb, View out, long long M, int C, cudaStream_t st) {
  launch(k_film, cdiv(M * (C / 8), 256), 256, 0, st, y, gb, out, M, C);
}

// F.adaptive_avg_pool2d to (H/16, W/16) + clamp(0,1)  (seg_video_model_fast.py:306-307)
__global__ void k_avgpool16_clamp(const float* mask, float* __restrict__ out, int B,
                                  int H, int W, const void* const* slot) {
  pdl_prologue_done();
  mask = io_ptr(mask, slot);
  int Wo = W / 16, Ho = H / 16;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  int wo = (int)(idx % Wo), ho = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  const float* src = mask + ((long long)b * H + ho * 16) * W + wo * 16;
  float s = 0.0f;
  for (int y = 0; y < 16; ++y) {
    const float4* r = reinterpret_cast<const float4*>(src + (long long)y * W);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      float4 v = r[x];
      s = add_rn(s, v.x); s = add_rn(s, v.y); s = add_rn(s, v.z); s = add_rn(s, v.w);
    }
  }
  s = s / 256.0f;
  out[idx] = fminf(fmaxf(s, 0.0f), 1.0f);
}
void avgpool16_clamp(const float* mask, float* out, int B, int H, int W, cudaStream_t st, const void* const* slot) {
  long long n = (long long)B * (H / 16) * (W / 16);
  launch(k_avgpool16_clamp, cdiv(n, 128), 128, 0, st, mask, out, B, H, W, slot);
}

// MaskFiLM: 3x3 (1->16) + ReLU + 1x1 (16->2C), then hyper_in = y*(1+gamma)+beta
// (seg_video_model_fast.py:159-180,312-314).  One block per pixel, one thread per channel.
// y / out live on an H x W grid; the mask map m is Hm x Wm <= H x W and counts as zero outside (the reference pads
// y by replication and the pooled mask with zeros, seg_video_model_fast.py:309-311)
__global__ void k_maskfilm_apply(const float* __restrict__ m, View y, View out,
                                 const float* __restrict__ w0, const float* __restrict__ b0,
                                 const float* __restrict__ w2, const float* __restrict__ b2, int H,
                                 int W, int C, int Hm, int Wm) {
  pdl_prologue_done();
  __shared__ float hid[16];
  long long pix = blockIdx.x;
  int w = (int)(pix % W), h = (int)((pix / W) % H);
  long long b = pix / ((long long)W * H);
  if (threadIdx.x < 16) {
    float acc = 0.0f;
    if (m) {
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          if ((unsigned)(h + dy) >= (unsigned)Hm || (unsigned)(w + dx) >= (unsigned)Wm) continue;
          acc = fmaf(m[(b * Hm + h + dy) * Wm + w + dx], w0[threadIdx.x * 9 + (dy + 1) * 3 + dx + 1], acc);
        }
    }
    hid[threadIdx.x] = fmaxf(add_rn(acc, b0[threadIdx.x]), 0.0f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g = 0.0f, b = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      g = fmaf(hid[k], w2[c * 16 + k], g);
      b = fmaf(hid[k], w2[(C + c) * 16 + k], b);
    }
    g = add_rn(g, b2[c]);
    b = add_rn(b, b2[C + c]);
    float v = ld3(y, pix, c);
    st3(out, pix, c, add_rn(mul_rn(v, add_rn(1.0f, g)), b));
  }
}
void maskfilm_apply(const float* m, View y, View out, const float* w0, const float* b0,
                    const float* w2, const float* b2, int B, int H, int W, int C, int Hm, int Wm, cudaStream_t st) {
  launch(k_maskfilm_apply, (unsigned)((long long)B * H * W), 128, 0, st, m, y, out, w0, b0, w2, b2, H, W, C,
                                                                                    Hm, Wm);
}

// F.interpolate(bilinear, align_corners=False) by exactly 1/8 and 8 (mask_predictor.py:35,44)
__global__ void k_bilinear_down8(const float* in, float* __restrict__ out, int B, int H,
                                 int W, const void* const* slot) {
  pdl_prologue_done();
  in = io_ptr(in, slot);
  int Ho = H / 8, Wo = W / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  int wo = (int)(idx % Wo), ho = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  const float* p = in + ((long long)b * H + ho * 8 + 3) * W + wo * 8 + 3;
  float top = add_rn(mul_rn(0.5f, p[0]), mul_rn(0.5f, p[1]));
  float bot = add_rn(mul_rn(0.5f, p[W]), mul_rn(0.5f, p[W + 1]));
  out[idx] = add_rn(mul_rn(0.5f, top), mul_rn(0.5f, bot));
}
void bilinear_down8(const float* in, float* out, int B, int H, int W, cudaStream_t st, const void* const* slot) {
  long long n = (long long)B * (H / 8) * (W / 8);
  launch(k_bilinear_down8, cdiv(n, 256), 256, 0, st, in, out, B, H, W, slot);
}
__global__ void k_bilinear_up8(const float* __restrict__ in, float* out, int B, int h,
                               int w, const void* const* slot) {
  pdl_prologue_done();
  out = io_ptr(out, slot);
  int Ho = h * 8, Wo = w * 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  float sy = fmaxf(sub_rn(mul_rn(0.125f, add_rn((float)oy, 0.5f)), 0.5f), 0.0f);
  float sx = fmaxf(sub_rn(mul_rn(0.125f, add_rn((float)ox, 0.5f)), 0.5f), 0.0f);
  int y0 = (int)sy, x0 = (int)sx;
  int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  float ly1 = sub_rn(sy, (float)y0), lx1 = sub_rn(sx, (float)x0);
  float ly0 = sub_rn(1.0f, ly1), lx0 = sub_rn(1.0f, lx1);
  const float* p = in + (long long)b * h * w;
  float top = add_rn(mul_rn(lx0, p[y0 * w + x0]), mul_rn(lx1, p[y0 * w + x1]));
  float bot = add_rn(mul_rn(lx0, p[y1 * w + x0]), mul_rn(lx1, p[y1 * w + x1]));
  out[idx] = add_rn(mul_rn(ly0, top), mul_rn(ly1, bot));
}
void bilinear_up8(const float* in, float* out, int B, int h, int w, cudaStream_t st, const void* const* slot) {
  long long n = (long long)B * h * 8 * w * 8;
  launch(k_bilinear_up8, cdiv(n, 256), 256, 0, st, in, out, B, h, w, slot);
}

__global__ void k_conv3x3_c1(const float* __restrict__ in, const float* __restrict__ wt,
                             const float* __restrict__ bias, View out, int B, int H, int W, int C) {
  pdl_prologue_done();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)B * H * W;
  if (idx >= M * C) return;
  long long m = idx / C;
  int c = (int)(idx % C);
  int w = (int)(m % W), h = (int)((m / W) % H);
  float acc = 0.0f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      if ((unsigned)(h + dy) >= (unsigned)H || (unsigned)(w + dx) >= (unsigned)W) continue;
      acc = fmaf(in[m + dy * W + dx], wt[c * 9 + (dy + 1) * 3 + dx + 1], acc);
    }
  st3(out, m, c, add_rn(acc, bias[c]));
}
void conv3x3_c1(const float* in, const float* w, const float* b, View out, int B, int h, int w_,
                int C, cudaStream_t st) {
  long long n = (long long)B * h * w_ * C;
  launch(k_conv3x3_c1, cdiv(n, 256), 256, 0, st, in, w, b, out, B, h, w_, C);
}

__global__ void k_conv1x1_to1(View in, const float* __restrict__ wt, const float* __restrict__ bias,
                              float* __restrict__ out, long long M, int K) {
  pdl_prologue_done();
  long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc = 0.0f;
  for (int k = lane; k < K; k += 32) acc = fmaf(ld3(in, m, k), wt[k], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[m] = add_rn(acc, bias[0]);
}
void conv1x1_to1(View in, const float* w, const float* b, float* out, long long M, int K,
                 cudaStream_t st) {
  launch(k_conv1x1_to1, cdiv(M * 32, 256), 256, 0, st, in, w, b, out, M, K);
}

// ------------------------------------------------------------------ caller-side statistics
// trainer_seg_video_model.py:655-660 (_roi_mse), :904-934 (mse), bits from bpp.
// grid = (chunks of a plane, B * 3 planes); 4 pixels per thread and step (HW is a multiple of 4: H, W multiples of 16)
__global__ void k_frame_stats(double* stats, const float* __restrict__ xh, const float* __restrict__ x,
                              const float* __restrict__ mask, long long HW, long long n) {
  pdl_prologue_done();
  (void)n;
  const long long plane = blockIdx.y;                 // b * 3 + c
  const float4* ph = reinterpret_cast<const float4*>(xh + plane * HW);
  const float4* px = reinterpret_cast<const float4*>(x + plane * HW);
  const float4* pm = mask ? reinterpret_cast<const float4*>(mask + (plane / 3) * HW) : nullptr;
  double se = 0.0, rse = 0.0, rn = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW / 4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 a = ph[i], t = px[i];
    const float d[4] = {sub_rn(a.x, t.x), sub_rn(a.y, t.y), sub_rn(a.z, t.z), sub_rn(a.w, t.w)};
    float mv[4] = {0.f, 0.f, 0.f, 0.f};
    if (pm) { const float4 m = pm[i]; mv[0] = m.x; mv[1] = m.y; mv[2] = m.z; mv[3] = m.w; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d2 = mul_rn(d[k], d[k]);
      se += (double)d2;
      if (mv[k] > 0.0f) { rse += (double)d2; rn += 1.0; }
    }
  }
  se = block_sum(se);
  rse = block_sum(rse);
  rn = block_sum(rn);
  if (threadIdx.x == 0) {
    atomicAdd(&stats[2], se);
    if (mask) { atomicAdd(&stats[3], rse); atomicAdd(&stats[4], rn); }
  }
}
__global__ void k_frame_stats_bits(double* stats, const float* bpp3, int B, double pixels, double n) {
  pdl_prologue_done();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (bpp3) {
      double by = 0, bz = 0;
      for (int b = 0; b < B; ++b) { by += (double)bpp3[3 * b + 1] * pixels; bz += (double)bpp3[3 * b + 2] * pixels; }
      stats[0] += by; stats[1] += bz;
    }
    stats[5] += n;
    stats[6] += (double)B;
  }
}
void frame_stats(double* stats7, const float* x_hat, const float* x, const float* mask,
                 const float* bpp3, int B, int H, int W, cudaStream_t st) {
  long long HW = (long long)H * W, n = (long long)B * 3 * HW;
  // (x_hat, x and mask are fp32 NCHW, 16-byte aligned planes: HW is a multiple of 4)
  unsigned gx = (unsigned)min((long long)(num_sms() * 8 / (3 * B) + 1), (long long)cdiv(HW / 4, 256));
  if (gx < 1) gx = 1;
  dim3 grid(gx, 3 * B);
  launch(k_frame_stats, grid, 256, 0, st, stats7, x_hat, x, mask, HW, n);
  launch(k_frame_stats_bits, 1, 32, 0, st, stats7, bpp3, B, (double)HW, (double)n);
}

// ---------------------------------------------------------------------------------------------------------------
// Device data path: decoded camera frames (uint8, interleaved) + cached masks (uint8) -> the
// (N, 4, h, w) fp32 [Y, Cb, Cr, mask] tensors the trainer feeds the codec, cropped on the way.
//   src/dataset/seg_waymo_dataset.py:26-34   rgb = uint8 / 255.0
//   src/dataset/seg_waymo_dataset.py:36-43   BT.709: y = Kr r + Kg g + Kb b; cb = 0.5 (b - y) / (1 - Kb) + 0.5; cr likewise; clamp
//   src/dataset/seg_waymo_dataset.py:56-79   mask in {0, 1} (npz: stored 0/1, png: > 127)
//   src/dataset/seg_waymo_dataset.py:231-245 one crop for the whole sequence, mask appended as channel 4
// Every operation in the reference's order with IEEE roundings (the reference runs these lines on the CPU in fp32: true
// divisions, no FMA contraction): the output is bit-identical.  4 pixels per thread: 12 + 4 bytes in, 4 x 16 bytes out.
constexpr int kFrameRows = 8;
__global__ void k_frames_from_u8(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, float* __restrict__ out,
                                 int H0, int W0, int top, int left, int h, int w, int out_ch, int bgr, int mask_thr) {
  pdl_prologue_done();
  // value / 255.0f for the 256 possible bytes, each computed once per block with the IEEE division the reference
  // performs per element (three divisions per pixel otherwise: the kernel was bound by instruction issue, 2.7 TB/s)
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = __fdiv_rn((float)i, 255.0f);
  __syncthreads();
  const int n = blockIdx.z;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= w) return;
  const size_t plane = (size_t)h * w;
  // (a block walks kFrameRows image rows: the table above is built once per 8 rows x 512 pixels)
  for (int y = blockIdx.y * kFrameRows; y < min(h, (int)(blockIdx.y + 1) * kFrameRows); ++y) {
  const uint8_t* src = img + (((size_t)n * H0 + (top + y)) * W0 + (left + x0)) * 3;
  const uint8_t* msrc = mask ? mask + ((size_t)n * H0 + (top + y)) * W0 + (left + x0) : nullptr;
  float* dst = out + (size_t)n * out_ch * plane + (size_t)y * w + x0;
  const int cnt = min(4, w - x0);
  const float Kr = 0.2126f, Kg = 0.7152f, Kb = 0.0722f;
  const float dcb = (float)(1.0 - 0.0722), dcr = (float)(1.0 - 0.2126);      // Python doubles, rounded when they meet the tensor
  // 12 image bytes and 4 mask bytes of this thread: three / one 32-bit loads where the addresses allow it
  uint8_t px[12], mk[4] = {0, 0, 0, 0};
  if (cnt == 4 && ((uintptr_t)src & 3) == 0) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint32_t v = __ldg(s32 + j);
      px[4 * j] = (uint8_t)v; px[4 * j + 1] = (uint8_t)(v >> 8); px[4 * j + 2] = (uint8_t)(v >> 16); px[4 * j + 3] = (uint8_t)(v >> 24);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 12; ++j) px[j] = j < 3 * cnt ? __ldg(src + j) : (uint8_t)0;
  }
  if (msrc) {
    if (cnt == 4 && ((uintptr_t)msrc & 3) == 0) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(msrc));
      mk[0] = (uint8_t)v; mk[1] = (uint8_t)(v >> 8); mk[2] = (uint8_t)(v >> 16); mk[3] = (uint8_t)(v >> 24);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) mk[j] = j < cnt ? __ldg(msrc + j) : (uint8_t)0;
    }
  }
  float Y[4], Cb[4], Cr[4], M[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float c0 = lut[px[3 * i]], c1 = lut[px[3 * i + 1]], c2 = lut[px[3 * i + 2]];
    const float r = bgr ? c2 : c0, g = c1, b = bgr ? c0 : c2;
    const float yy = __fadd_rn(__fadd_rn(__fmul_rn(Kr, r), __fmul_rn(Kg, g)), __fmul_rn(Kb, b));
    const float cb = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(b, yy)), dcb), 0.5f);
    const float cr = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(r, yy)), dcr), 0.5f);
    Y[i] = fminf(fmaxf(yy, 0.0f), 1.0f);
    Cb[i] = fminf(fmaxf(cb, 0.0f), 1.0f);
    Cr[i] = fminf(fmaxf(cr, 0.0f), 1.0f);
    M[i] = (int)mk[i] > mask_thr ? 1.0f : 0.0f;
  }
  const bool vec = cnt == 4 && (w & 3) == 0;                 // (then every row start is 16-byte aligned)
  float* planes[4] = {dst, dst + plane, dst + 2 * plane, dst + 3 * plane};
  const float* vals[4] = {Y, Cb, Cr, M};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= out_ch) break;
    if (vec) {
      *reinterpret_cast<float4*>(planes[c]) = make_float4(vals[c][0], vals[c][1], vals[c][2], vals[c][3]);
    } else {
      for (int i = 0; i < cnt; ++i) planes[c][i] = vals[c][i];
    }
  }
  }
}
void frames_from_u8(const uint8_t* img, const uint8_t* mask, float* out, int N, int H0, int W0, int top, int left, int h,
                    int w, int out_ch, int bgr, int mask_thr, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(cdiv(w, 4), 128), (unsigned)cdiv(h, kFrameRows), (unsigned)N);
  launch(k_frames_from_u8, grid, 128, 0, st, img, mask, out, H0, W0, top, left, h, w, out_ch, bgr, mask_thr);
}

// Mask propagation: the predictor's logits of frame t-1, thresholded at sigma = 0.5
// (logit > 0), are the mask frame t is coded with.
__global__ void k_mask_from_logits(const float* __restrict__ logits, float* __restrict__ mask, long long n) {
  pdl_prologue_done();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(logits + i);
    *reinterpret_cast<float4*>(mask + i) = make_float4(v.x > 0.f ? 1.f : 0.f, v.y > 0.f ? 1.f : 0.f, v.z > 0.f ? 1.f : 0.f,
                                                       v.w > 0.f ? 1.f : 0.f);
  } else {
    for (long long j = i; j < n; ++j) mask[j] = logits[j] > 0.f ? 1.f : 0.f;
  }
}
void mask_from_logits(const float* logits, float* mask, long long n, cudaStream_t st) {
  launch(k_mask_from_logits, (unsigned)cdiv(cdiv(n, 4LL), 256LL), 256, 0, st, logits, mask, n);
}

}  // namespace dmc
