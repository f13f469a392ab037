// Host-side launchers of every kernel of the engine.  All launch asynchronously on `st`.
#pragma once
#include "common.cuh"
#include <string.h>

namespace dmc {

int num_sms();
void note_launch();            // every kernel launch of the library is counted
long long launch_count();
void add_launches(long long n);   // kernels replayed from a captured graph
// accumulate-truncation compensation of the tcgen05 kernels (kernels.cu): kappa in units of 2^-24 per MMA step
float acc_comp_kappa();
void acc_comp_set_kappa(float k);
float acc_comp_scaled(int K);      // kappa * (K/16 + 1) * 2^-24 * 2^11 (the scale of the second accumulator)
bool pdl_enabled();            // programmatic dependent launch: DMC_PDL=0/1 forces it, else per frame size
void pdl_set_auto(bool on);

// Launch with the programmatic-stream-serialization attribute (see pdl_prologue_done in common.cuh).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  note_launch();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------- weights ----------------
// Packed contraction weight: split planes [2][Npad][Kld] fp16 (hi, 2^11-scaled lo), K index = (kh*KW + kw)*Cin + ci.
struct GemmW {
  h16* w = nullptr;
  h16* wb = nullptr;       // tile-blocked, pre-swizzled copy [2][Kld/16][Npad][16] for the tcgen05 kernels (or nullptr)
  float* bias = nullptr;   // [Npad], packed order, zero on padding
  int N = 0, K = 0;        // logical (N = cout, K = cin*kh*kw)
  int ncols = 0;           // packed columns before tile padding (pair / shuffle layouts included)
  int Npad = 0, Kld = 0;   // allocated rows (multiple of 64 and of BN) / row pitch (multiple of 64)
  int pack = PACK_PLAIN;
  int BN = 128;            // tcgen05 N tile this weight was padded for
  int Cg = 0, Cg_pad = 0;  // PACK_SHUF2
  void* tmap = nullptr;    // host CUtensorMap (128 B), box = 64 k x BN rows (one-CTA kernel)
  void* tmap_half = nullptr;  // box = 64 k x BN/2 rows (CTA-pair kernel: each CTA loads half of W)
  void* tmap_s3 = nullptr;    // box = 32 k x BN/2 rows x 2 planes (gemm_s3.cu)
  void* tmap_s3_hi = nullptr; // same with the hi plane only (single-term products)
};
// transposed (1x1): w is the (cin, cout, 1, 1) weight of another convolution, packed as its transpose
void pack_gemm_weight(const float* w_oihw, int cout, int cin, int kh, int kw, const GemmW& g,
                      cudaStream_t st, bool transposed = false);
void pack_gemm_bias(const float* bias, int cout, const GemmW& g, cudaStream_t st);   // nullptr -> zeros
// depthwise 3x3: (C,1,3,3) -> [9][C] fp32
void pack_dw_weight(const float* w, float* out9c, int C, cudaStream_t st);

// ---------------- caller-owned tensors ----------------
// Device-resident table of the tensors of the current forward call.  Kernels that read / write them take an optional
// `slot` (address of one entry): the pointer is then fetched at run time, so a captured CUDA graph is independent of
// the caller's buffers.
struct IoSlots {
  const void* x; const void* mask; const void* dpb_frame; const void* dpb_feature;
  const void* x_hat; const void* feature; const void* bpp3; const void* mask_pred; const void* mask_src; const void* finite;
};
void set_io(IoSlots* dst, const IoSlots& v, cudaStream_t st);
void copy_flag(const int* src, const void* const* slot, cudaStream_t st);

// ---------------- layout conversion ----------------
// (B,Cimg,H,W) fp32 NCHW -> S3 [B*H/8*W/8, Cimg*64], channel = c*64 + dy*8 + dx (F.pixel_unshuffle)
void unshuffle8_in(const float* x, View out, int B, int Cimg, int H, int W, cudaStream_t st, const void* const* slot = nullptr);
// fp32 row-major [M, Cimg*64] (ld) -> (B,Cimg,H,W) NCHW, clamped to [0,1] (F.pixel_shuffle + clamp)
void shuffle8_out(const float* in, int ld, float* x, int B, int Cimg, int H, int W, cudaStream_t st, const void* const* slot = nullptr);
void nchw_to_s3(const float* x, View out, int B, int C, int H, int W, cudaStream_t st, const void* const* slot = nullptr);
void s3_to_nchw(View in, float* x, int B, int C, int H, int W, cudaStream_t st, const void* const* slot = nullptr);
void f32rows_to_nchw(const float* in, int ld, float* x, int B, int C, int H, int W, cudaStream_t st);
void scale_cols(View in, const float* scale, View out, long long M, cudaStream_t st);
void copy_view(View in, View out, long long M, cudaStream_t st);
// (B,Hin,Win,C) -> (B,Hout,Wout,C): replicate pad right/bottom (inference.py:40-43) or crop to the top-left corner
void regrid(View in, int Hin, int Win, View out, int Hout, int Wout, int B, cudaStream_t st);
// bit i of *flag |= tensor i has a non-finite / fp16-saturated element (one launch for up to 8 tensors)
struct FiniteList { View v[8]; long long M[8]; int n; };
void finite_check(const FiniteList& l, int* flag, cudaStream_t st);

// ---------------- convolution pieces ----------------
void dwconv3x3(View in, const float* w9c, const float* bias, View out, int B, int H, int W,
               cudaStream_t st);
// same with the input as fp32 rows [M, ld] (what the contraction in front of it writes)
void dwconv3x3_f32(const float* in, int ld, const float* w9c, const float* bias, View out, int B, int H, int W,
                   cudaStream_t st);
// dst columns: tap * tap_stride + col_off + c
void im2col(View in, View out, int B, int H, int W, int k, int stride, int pad, int Ho, int Wo,
            int tap_stride, int col_off, cudaStream_t st);
void gemm_simt(View a, const GemmW& w, const Epi& e, long long M, cudaStream_t st);

// tcgen05 path (gemm_umma.cu).  `tmapA` is a host CUtensorMap built by make_tmap_act.
int make_tmap_act(void* tmap_out, View a, long long M);                    // box 64 x 128
int make_tmap_weight(void* tmap_out, const GemmW& w, int box_rows);        // box 64 x box_rows
void umma_set_debug(int mask);  // probe switches of k_gemm_umma (bench tool only)
void umma_set_pair(bool on);   // use the cta_group::2 kernel where possible (default on)
int gemm_umma(const void* tmapA, const GemmW& w, const Epi& e, long long M, int K, int nsplit,
              cudaStream_t st);
const char* umma_last_error();

// Specialised CTA-pair kernel with a TMA epilogue (gemm_s3.cu): S3 out, plain / chunk-add-pair column
// layouts, <= 1 residual, 3-term split-fp16 product.  gemm_s3_supports() says whether a launch qualifies.
int make_tmap_s3_act(void* tmap_out, View a, long long M, int planes);     // box 32 x 128 x planes
int make_tmap_s3_act64(void* tmap_out, View a, long long M);               // box 32 x 64 x 1
int make_tmap_s3_weight(void* tmap_out, const GemmW& w, int planes);       // box 32 x BN/2 x planes
int make_tmap_s3_rows(void* tmap_out, View v, int cols, long long M);      // box 16 x 32 x 2 planes
int make_tmap_f32_rows(void* tmap_out, float* base, int cols, int ld, long long M);   // fp32 rows, box 16 x 32
bool gemm_s3_supports(const GemmW& w, const Epi& e, int nsplit);
void gemm_s3_set_debug(int mask);
void gemm_s3_set_plain_launch(int on);   // measurement passes: no cooperative launch attribute (gemm_s3.cu)
// A chain = consecutive 1x1 layers over the same rows, run by ONE persistent launch with tile-level
// dependencies (see gemm_s3.cu).  A single contraction is a chain of one.
struct S3StageDesc {
  const void* tmA;            // host CUtensorMap of the A operand (make_tmap_s3_act)
  const void* tmA64;          // the same operand as 64-row single-plane boxes (make_tmap_s3_act64) or nullptr
  const GemmW* w;
  Epi e;                      // scale is resolved at launch from scale_table / scale_C and qp
  const void* tmOut;          // make_tmap_s3_rows or make_tmap_f32_rows
  const void* tmRes;          // residual (make_tmap_s3_rows) or nullptr
  const void* tmRes2;         // second residual (shortcut blocks) or nullptr
  int K, nsplit;
  const float* scale_table;   // (72, scale_C) per-QP table or nullptr
  int scale_C;
};
struct S3Chain;
int s3_chain_max_stages();
S3Chain* s3_chain_create(const S3StageDesc* stages, int n, long long M);   // nullptr on error
void s3_chain_destroy(S3Chain* c);
int s3_chain_stages(const S3Chain* c);
int s3_chain_launch(S3Chain* c, int qp, cudaStream_t st);
const char* gemm_s3_last_error();
int gemm_s3_trap_code();     // which bounded wait of the chain kernel gave up (0 = none), readable after a trap

// ---------------- entropy model ----------------
struct PriorArgs {
  int scheme;          // 2 or 4 (checkerboard steps)
  int step;
  int B, H, W, C;
  View y;              // latent y (after FiLM for `performance`)
  View params;         // scheme 2: [q_dec | sigma0 | mu0] (3C); scheme 4: [qe, qd | sigma0 | mu0] (2+2C)
  View sp;             // steps >= 1: spatial prior output [sigma | mu] (2C)
  View yh;             // running y_hat (pre q_dec), C columns, written for owned elements
  float* sym;          // [M, C] fp32
  float* sig;          // [M, C] fp32
  int mode;            // 0 encoder: symbols = round(y / q - mu).  Decoder side (dmc_decode_*): 2 = only publish the
                       // sigmas of this step's elements (what the entropy decoder needs first), 1 = take this step's
                       // symbols from sym_in and rebuild y_hat from them
  const float* sym_in; // mode 1: dense (B, C, H, W) fp32, read at this step's elements only
};
void prior_step(const PriorArgs& a, cudaStream_t st);
// y_hat = yh * q_dec -> out; bits(sym, sig) summed per sample into bits_acc[b] (double)
void prior_finish(const PriorArgs& a, View y_hat, int formula, double* bits_acc, cudaStream_t st);
// z -> round -> z_hat (S3) + factorized bits; tables: 11 pointers to rows (C floats) for this qp:
// f1.h f1.b f1.a f2.h f2.b f2.a f3.h f3.b f3.a f4.h f4.b
struct BitparmRow { const float* p[11]; };
void round_z_bits(View z, View z_hat, int B, int HW, int C, BitparmRow t, double* bits_acc,
                  cudaStream_t st);
void finalize_bpp(const double* bits_y, const double* bits_z, float* bpp3, int B, int pixels,
                  cudaStream_t st, const void* const* slot = nullptr);
void gaussian_bits(const float* sym, const float* sigma, float* bits, long long n, int formula,
                   cudaStream_t st);

// ---------------- mask conditioning ----------------
void film(View y, View gb, View out, long long M, int C, cudaStream_t st);   // y*(1+g)+b, gb=[g|b]
// 16x16 block mean of an fp32 (B,1,H,W) map, clamped to [0,1] -> (B,H/16,W/16) fp32
void avgpool16_clamp(const float* mask, float* out, int B, int H, int W, cudaStream_t st, const void* const* slot = nullptr);
// MaskFiLM (3x3 1->16, ReLU, 1x1 16->2C) + FiLM on y;  m == nullptr means an all-zero mask; m is an Hm x Wm map
// (zero outside) on y's H x W grid
void maskfilm_apply(const float* m, View y, View out, const float* w0, const float* b0,
                    const float* w2, const float* b2, int B, int H, int W, int C, int Hm, int Wm, cudaStream_t st);
void bilinear_down8(const float* in, float* out, int B, int H, int W, cudaStream_t st, const void* const* slot = nullptr);  // -> H/8
void bilinear_up8(const float* in, float* out, int B, int h, int w, cudaStream_t st, const void* const* slot = nullptr);    // -> 8h
// 3x3 conv with a single input channel (mask_embed): fp32 map (B,h,w) -> S3 [M, C]
void conv3x3_c1(const float* in, const float* w, const float* b, View out, int B, int h, int w_,
                int C, cudaStream_t st);
// 1x1 conv to a single channel: S3 [M,K] -> fp32 [M]
void conv1x1_to1(View in, const float* w, const float* b, float* out, long long M, int K,
                 cudaStream_t st);

// ---------------- caller-side statistics ----------------
void frame_stats(double* stats7, const float* x_hat, const float* x, const float* mask,
                 const float* bpp3, int B, int H, int W, cudaStream_t st);
// uint8 camera frames (N, H0, W0, 3) + uint8 masks (N, H0, W0) or NULL -> (N, out_ch, h, w) fp32 [Y, Cb, Cr(, mask)],
// cropped at (top, left); seg_waymo_dataset.py:26-43,56-79,231-245
void frames_from_u8(const uint8_t* img, const uint8_t* mask, float* out, int N, int H0, int W0, int top, int left, int h,
                    int w, int out_ch, int bgr, int mask_thr, cudaStream_t st);
// mask = (logits > 0) as fp32 {0, 1}
void mask_from_logits(const float* logits, float* mask, long long n, cudaStream_t st);

// ---------------- training mode (train.cu) ----------------
void wsilu_rows(const float* in, float* out, long long n, cudaStream_t st);                       // fp32 rows, n % 4 == 0
// out = g * wsilu'(pre) + per-block column sums of out; returns the number of partial rows
int wsilu_bwd_parts(long long M, int C, int max_parts);
int wsilu_bwd(View g, const float* pre, int ld, View out, long long M, float* part, int ldp, int max_parts, cudaStream_t st);
// forward value v, pre-activation gradient gu and gu's per-block column sums in one pass; returns the number of partial rows
int chunkadd_parts(long long M, int C2);      // partial rows chunkadd_fwd_bwd writes (size `part` for them)
int chunkadd_fwd_bwd(const float* u, int ld, View gv, View v, View gu, long long M, float* part, int ldp, int max_parts,
                     cudaStream_t st);
// column sums of g (times h, element by element, when h != nullptr) as per-block partial rows; returns their number
int colsum_s3(View g, const View* h, long long M, float* part, int ldp, int max_parts, cudaStream_t st);
void reduce_partials(const float* part, long long stride, int S, float* out, long long n, const float* scale_dev,
                     float scale, cudaStream_t st);
// depthwise 3x3 weight (C,1,3,3) + bias gradient as partial rows of C * 10 floats ([c][tap], tap 9 = bias)
int dw_wgrad(const float* g, int ldg, int C, const float* t, int ld, int B, int H, int W, float* part, int ldp, int max_parts,
             cudaStream_t st);
void flip_dw_weight(const float* w9c, float* out, int C, cudaStream_t st);
void reduce_dw(const float* part, long long stride, int S, float* gw, float* gb, int C, const float* scale_dev, float scale,
               cudaStream_t st);
// scale2 = {2^floor(peak_log2 - log2 max|g|), its reciprocal}; part: 8 * num_sms() floats of scratch
void grad_scale(const float* g, long long n, float peak_log2, float* part, float* scale2, cudaStream_t st);
void nchw_to_s3_scaled(const float* x, View out, int B, int C, int H, int W, const float* chan, const float* gscale,
                       cudaStream_t st);
void s3_to_nchw_scaled(View in, float* x, int B, int C, int H, int W, const float* gscale, cudaStream_t st);
int nchw_dot(const float* a, const float* b, int B, int C, long long HW, float* part, int max_parts, cudaStream_t st);
void reduce_div(const float* part, int S, int C, const float* div, float* out, cudaStream_t st);
// dW[n][k] = sum_m G[m][n] X[m][k] as wgrad_splits() partial matrices of N * K floats (terms: 3 = fp32-grade split product)
int wgrad_splits(long long M, int N, int K);
// tcgen05 version (wgrad_umma.cu); wgrad_s3 routes to it unless DMC_WGRAD_UMMA=0
bool wgrad_umma_supported(View G, View X);
int wgrad_umma_splits(long long M, int N, int K);
int wgrad_umma(View G, View X, long long M, int terms, float* part, cudaStream_t st);   // -1 on error
const char* wgrad_umma_last_error();
int wgrad_s3(View G, View X, long long M, int terms, float* part, cudaStream_t st);
// k x k convolutions through the im2col view (train.cu)
void wt_kxk(const float* w, float* wt, int cout, int cin, int k, cudaStream_t st);
void reduce_wgrad_kxk(const float* part, long long stride, int S, float* out, int cout, int cin, int k, const float* scale_dev,
                      cudaStream_t st);
void col2im_nchw(View gcol, float* gx, int B, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo, const float* gscale,
                 cudaStream_t st);
void quant_train(const float* x, const float* noise, float* out, long long n, int mode, cudaStream_t st);
void gaussian_bits_bwd(const float* sym, const float* sigma, const float* go, float* gsym, float* gsig, long long n,
                       int formula, cudaStream_t st);

}  // namespace dmc
