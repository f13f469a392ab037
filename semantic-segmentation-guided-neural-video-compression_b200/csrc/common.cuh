// Shared types of the DMC engine: the split-bf16 ("S3") activation format, the GEMM
// epilogue description and the exact-arithmetic helpers.
//
// S3 format.  Every activation lives in HBM as THREE bf16 planes hi/mid/lo (row = pixel, column =
// channel) with hi+mid+lo == the fp32 value exactly (8+8+8 mantissa bits).
//
// Layout of a plane ("tile-blocked"): [C/16 column blocks][Mp rows][16 columns], Mp = rows padded to 256,
// and inside a row of a block the two 16-byte halves are swapped where bit 2 of the row index is set.
// That is byte for byte the shared-memory image of a K-major SWIZZLE_32B tcgen05 operand, so a
// 128-row x 32-column operand tile is two contiguous 4 KB pieces and a 32-row x 16-column epilogue
// chunk is one contiguous 1 KB piece: TMA moves them as 512-byte segments.  (The TMA unit retires
// only ~0.55 row segments per clock per SM whatever their width; with row-major planes a tile was
// ~10 700 segments of 32-64 bytes = 17 000 clocks of TMA work against 6 100 clocks of MMAs.)  The planes are what tcgen05 consumes directly (TMA -> swizzled smem -> kind::f16
// MMA), so a contraction at fp32-grade accuracy is 6 bf16 MMA terms (hh,hm,mh,hl,lh,mm)
// accumulated in fp32 TMEM, and the same buffer read with only the hi plane is a plain
// bf16 GEMM operand.  SURVEY.md 7.1: the reference's symbol-parity gate needs >= ~20
// operand mantissa bits, which plain bf16 / TF32 / 2-term splits do not give.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmc {

typedef __nv_bfloat16 bf16;

// A (possibly column-sliced, at multiples of 16 columns) view of an S3 tensor: the 16-byte unit
// holding columns [c, c+8) of row r in plane pl is at
//   p + pl * ps + (c >> 4) * bs + r * 16 + ((((c >> 3) & 1) ^ ((r >> 2) & 1)) << 3)        (elements)
struct View {
  bf16* p;
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

__device__ __forceinline__ void split3(float x, bf16& h, bf16& m, bf16& l) {
  h = __float2bfloat16_rn(x);
  float r = sub_rn(x, __bfloat162float(h));
  m = __float2bfloat16_rn(r);
  r = sub_rn(r, __bfloat162float(m));
  l = __float2bfloat16_rn(r);
}
__device__ __forceinline__ float join3(bf16 h, bf16 m, bf16 l) {
  return add_rn(add_rn(__bfloat162float(l), __bfloat162float(m)), __bfloat162float(h));
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(bf16 a, bf16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__device__ __forceinline__ float ld3(const View& v, long long row, int col) {
  const bf16* q = v.p + s3_unit_offset(v, row, col) + (col & 7);
  return join3(q[0], q[v.ps], q[2 * v.ps]);
}
__device__ __forceinline__ void st3(const View& v, long long row, int col, float x) {
  bf16 h, m, l;
  split3(x, h, m, l);
  bf16* q = v.p + s3_unit_offset(v, row, col) + (col & 7);
  q[0] = h;
  q[v.ps] = m;
  q[2 * v.ps] = l;
}

// 8 consecutive columns (16 B per plane); col must be a multiple of 8 and the view 16B aligned.
__device__ __forceinline__ void ld3x8(const View& v, long long row, int col, float* o) {
  const bf16* q = v.p + s3_unit_offset(v, row, col);
  uint4 a = *reinterpret_cast<const uint4*>(q);
  uint4 b = *reinterpret_cast<const uint4*>(q + v.ps);
  uint4 c = *reinterpret_cast<const uint4*>(q + 2 * v.ps);
  const uint32_t* ua = &a.x;
  const uint32_t* ub = &b.x;
  const uint32_t* uc = &c.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = add_rn(add_rn(bf16lo(uc[i]), bf16lo(ub[i])), bf16lo(ua[i]));
    o[2 * i + 1] = add_rn(add_rn(bf16hi(uc[i]), bf16hi(ub[i])), bf16hi(ua[i]));
  }
}
__device__ __forceinline__ void st3x8(const View& v, long long row, int col, const float* x) {
  uint32_t a[4], b[4], c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bf16 h0, m0, l0, h1, m1, l1;
    split3(x[2 * i], h0, m0, l0);
    split3(x[2 * i + 1], h1, m1, l1);
    a[i] = pack_bf16(h0, h1);
    b[i] = pack_bf16(m0, m1);
    c[i] = pack_bf16(l0, l1);
  }
  bf16* q = v.p + s3_unit_offset(v, row, col);
  *reinterpret_cast<uint4*>(q) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(q + v.ps) = make_uint4(b[0], b[1], b[2], b[3]);
  *reinterpret_cast<uint4*>(q + 2 * v.ps) = make_uint4(c[0], c[1], c[2], c[3]);
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

}  // namespace dmc
