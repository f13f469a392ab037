// Shared types of the DMC engine: the split-fp16 ("S3" in identifiers: split storage) activation
// format, the GEMM epilogue description and the exact-arithmetic helpers.
//
// Split format.  Every activation lives in HBM as TWO fp16 planes (row = pixel, column = channel):
//   hi = fp16(x)                      (11 significand bits)
//   lo = fp16((x - hi) * 2^11)        (the next 11 bits; x - hi is exact in fp32)
// so  x ~= hi + lo * 2^-11  to 2^-23..2^-22 relative -- one fp32 rounding -- at 4 B/element.  The 2^11
// pre-scale keeps the low part in fp16's normal range (unscaled it goes subnormal and the symbol-parity
// gate fails).  A contraction at fp32-grade accuracy is THREE fp16 MMA terms per k step:
//   acc_main  += a_hi . w_hi
//   acc_small += a_hi . w_lo + a_lo . w_hi          (2^11-scaled; a_lo . w_lo <= 2^-22 is dropped)
//   result     = acc_main + acc_small * 2^-11
// accumulated in fp32 TMEM (0 symbol mismatches over three free-running P frames; plain
// bf16 / TF32 / unscaled 2-term splits fail the gate, the 6-term bf16x3 split of the first versions of this
// engine passed at twice the MMAs and 1.5x the operand bytes).  The same buffer read with only the hi
// plane is a plain fp16 GEMM operand (recon_generation_net).
// Range: |x| must stay below 65504 (fp16); the conversions saturate instead of producing inf.  Random-init
// activations peak at 4.6; the reference's own guard trips at 1e6 (trainer:285-287).
//
// Layout of a plane ("tile-blocked"): [C/16 column blocks][Mp rows][16 columns], Mp = rows padded to 256,
// and inside a row of a block the two 16-byte halves are swapped where bit 2 of the row index is set.
// That is byte for byte the shared-memory image of a K-major SWIZZLE_32B tcgen05 operand, so a
// 128-row x 32-column operand tile is two contiguous 4 KB pieces and a 32-row x 16-column epilogue
// chunk is one contiguous 1 KB piece: TMA moves them as 512-byte segments.  (The TMA unit retires
// only ~0.55 row segments per clock per SM whatever their width; with row-major planes a tile was
// ~10 700 segments of 32-64 bytes = 17 000 clocks of TMA work against 6 100 clocks of MMAs.)  The planes
// are what tcgen05 consumes directly (TMA -> swizzled smem -> kind::f16 MMA).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmc {

typedef __half h16;

constexpr int kPlanes = 2;                   // hi, lo
constexpr float kLoScale = 2048.0f;          // 2^11
constexpr float kLoInv = 1.0f / 2048.0f;

// A (possibly column-sliced, at multiples of 16 columns) view of an S3 tensor: the 16-byte unit
// holding columns [c, c+8) of row r in plane pl is at
//   p + pl * ps + (c >> 4) * bs + r * 16 + ((((c >> 3) & 1) ^ ((r >> 2) & 1)) << 3)        (elements)
struct View {
  h16* p;
  long long ps;   // plane stride, elements
  long long bs;   // column-block stride, elements (= padded rows * 16)
  int C;          // columns in this view
};
__host__ __device__ __forceinline__ long long s3_unit_offset(const View& v, long long row, int col) {
  return (long long)(col >> 4) * v.bs + row * 16 + ((((col >> 3) & 1) ^ (int)((row >> 2) & 1)) << 3);
}

enum { ACT_NONE = 0, ACT_WSILU = 1, ACT_RELU = 2 };
enum { PACK_PLAIN = 0, PACK_PAIR = 1, PACK_SHUF2 = 2 };

// What happens to one accumulator element after the contraction (layers.py:65-79 is the
// longest chain):  v = acc + bias; v = act(v); [pair: v = v + act(partner)];
// v += res1; v += res2; v *= scale[col]; clamp; store (S3 or fp32).
struct Epi {
  const float* bias;     // packed column order, length = packed N
  int act;
  int pack;              // PACK_* : how packed columns map to destination columns
  View res1, res2;       // p == nullptr -> absent; indexed by destination (row, col)
  const float* scale;    // per destination column or nullptr
  View out;              // S3 destination (p may be nullptr when out_f32 is used)
  float* out_f32;        // fp32 row-major destination or nullptr
  int ld_f32;
  int n_out;             // valid destination columns
  int H, W;              // PACK_SHUF2: source spatial size (rows = B*H*W)
  int Cg, Cg_pad;        // PACK_SHUF2: channels per pixel-shuffle group / padded
  int do_clamp;
  float clamp_lo, clamp_hi;
};

// ---- arithmetic that must round exactly like the reference's separate torch ops ----
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// layers.py:8-10  F.silu(4.0 * x) / 4.0
__device__ __forceinline__ float wsilu(float x) {
  float v = mul_rn(4.0f, x);
  float s = v / (1.0f + expf(-v));
  return mul_rn(s, 0.25f);
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_WSILU) return wsilu(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// two floats -> packed f16x2 (low half = a), round to nearest even, saturating at +-65504
__device__ __forceinline__ uint32_t cvt_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float h2lo(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu))); }
__device__ __forceinline__ float h2hi(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }

// x -> (hi, lo) of two elements at once: h = f16x2(x), l = f16x2((x - h) * 2^11)
__device__ __forceinline__ void split2x2(float x0, float x1, uint32_t& h, uint32_t& l) {
  h = cvt_h2(x0, x1);
  const float r0 = sub_rn(x0, h2lo(h)), r1 = sub_rn(x1, h2hi(h));
  l = cvt_h2(mul_rn(r0, kLoScale), mul_rn(r1, kLoScale));
}
__device__ __forceinline__ void split2(float x, h16& h, h16& l) {
  uint32_t a, b;
  split2x2(x, 0.0f, a, b);
  h = __ushort_as_half((unsigned short)(a & 0xffffu));
  l = __ushort_as_half((unsigned short)(b & 0xffffu));
}
// hi + lo * 2^-11 with one rounding (the product is exact)
__device__ __forceinline__ float join2(float h, float l) { return fmaf(l, kLoInv, h); }
__device__ __forceinline__ float join2(h16 h, h16 l) { return join2(__half2float(h), __half2float(l)); }

__device__ __forceinline__ float ld3(const View& v, long long row, int col) {
  const h16* q = v.p + s3_unit_offset(v, row, col) + (col & 7);
  return join2(q[0], q[v.ps]);
}
__device__ __forceinline__ void st3(const View& v, long long row, int col, float x) {
  h16 h, l;
  split2(x, h, l);
  h16* q = v.p + s3_unit_offset(v, row, col) + (col & 7);
  q[0] = h;
  q[v.ps] = l;
}

// 8 consecutive columns (16 B per plane); col must be a multiple of 8 and the view 16B aligned.
__device__ __forceinline__ void ld3x8(const View& v, long long row, int col, float* o) {
  const h16* q = v.p + s3_unit_offset(v, row, col);
  uint4 a = *reinterpret_cast<const uint4*>(q);
  uint4 b = *reinterpret_cast<const uint4*>(q + v.ps);
  const uint32_t* ua = &a.x;
  const uint32_t* ub = &b.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = join2(h2lo(ua[i]), h2lo(ub[i]));
    o[2 * i + 1] = join2(h2hi(ua[i]), h2hi(ub[i]));
  }
}
__device__ __forceinline__ void st3x8(const View& v, long long row, int col, const float* x) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2x2(x[2 * i], x[2 * i + 1], a[i], b[i]);
  h16* q = v.p + s3_unit_offset(v, row, col);
  *reinterpret_cast<uint4*>(q) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(q + v.ps) = make_uint4(b[0], b[1], b[2], b[3]);
}

// Packed GEMM column -> destination (row, col).  Returns false for padding columns.
// PACK_PAIR: groups of 64 packed columns = 32 channels c followed by their chunk-add
//            partners c + n_out (layers.py:12-20); only the first 32 yield an output.
// PACK_SHUF2: packed column = g * Cg_pad + c with g = dy*2+dx (nn.PixelShuffle(2)).
__device__ __forceinline__ bool epi_dest(const Epi& e, long long m, int n, long long& drow, int& dcol) {
  if (e.pack == PACK_PAIR) {
    if ((n & 63) >= 32) return false;
    dcol = (n >> 6) * 32 + (n & 31);
    drow = m;
  } else if (e.pack == PACK_SHUF2) {
    int g = n / e.Cg_pad;
    dcol = n - g * e.Cg_pad;
    if (dcol >= e.Cg) return false;
    int w = (int)(m % e.W);
    long long t = m / e.W;
    int h = (int)(t % e.H);
    long long b = t / e.H;
    drow = (b * (2 * e.H) + (2 * h + (g >> 1))) * (2LL * e.W) + (2 * w + (g & 1));
    return true;
  } else {
    dcol = n;
    drow = m;
  }
  return dcol < e.n_out;
}

// Scalar epilogue (SIMT GEMM and the reference for the vectorised tcgen05 epilogue).
// `acc2` is the partner accumulator (packed column n + 32) in PACK_PAIR mode.
__device__ __forceinline__ void epi_store(const Epi& e, long long m, int n, float acc, float acc2) {
  long long drow;
  int dcol;
  if (!epi_dest(e, m, n, drow, dcol)) return;
  float v = acc;
  if (e.bias) v = add_rn(v, e.bias[n]);
  v = apply_act(v, e.act);
  if (e.pack == PACK_PAIR) {
    float u = acc2;
    if (e.bias) u = add_rn(u, e.bias[n + 32]);
    v = add_rn(v, apply_act(u, e.act));
  }
  if (e.res1.p) v = add_rn(v, ld3(e.res1, drow, dcol));
  if (e.res2.p) v = add_rn(v, ld3(e.res2, drow, dcol));
  if (e.scale) v = mul_rn(v, e.scale[dcol]);
  if (e.do_clamp) v = fminf(fmaxf(v, e.clamp_lo), e.clamp_hi);
  if (e.out_f32) e.out_f32[drow * e.ld_f32 + dcol] = v;
  if (e.out.p) st3(e.out, drow, dcol, v);
}

// ---- programmatic dependent launch (small frames; DMC_PDL=0/1 forces it).  Every kernel of the frame is then launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_prologue_done(): its CTAs are scheduled
// as soon as the previous kernel's CTAs leave their SMs (that kernel released its dependents at its own start),
// run their prologue (barrier initialisation, TMEM allocation, index arithmetic), and block in griddepcontrol.wait
// until the previous grid has completed and its writes are visible.  No global memory is touched before the wait.
// Without the attribute both instructions are no-ops.  Measured on the 1920x1280 frame (power-capped box, three
// alternating runs each): 9 225 k clocks per frame with, 9 229 k without -- the persistent kernels hold every SM
// until their last tile, there is little tail to overlap.  On small frames, where the ~100 launches of a forward
// are bound by their fixed cost, it is worth 11 % (128x192: 1.61 -> 1.45 ms), so the engine turns it on below
// 1 Mpixel per forward.
__device__ __forceinline__ void pdl_prologue_done() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

}  // namespace dmc
