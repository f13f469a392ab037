/*
 * dmc_b200.h -- C ABI of the B200-native DMC P-frame / DMCI forward engine.
 *
 * This is the drop-in boundary for the one hot path this repository owns: the
 * `forward()` of the reference's codec modules.  Every entry point names the
 * reference interface it replaces (paths relative to the reference root):
 *
 *   dmc_forward   <-> DMC.forward(x, qp, dpb, after_i)
 *                       old          src/models/video_model.py:338-388
 *                       performance  src/refactor/seg_video_model.py:301-365
 *                       fast         src/refactor/seg_video_model_fast.py:328-411
 *                       mask_prop    src/refactor/mask_prop_seg_video_model.py:331-417
 *   dmci_forward  <-> DMCI.forward(x, qp)           src/models/image_model.py:205-261
 *   dmc_set_weight / dmc_finalize_weights
 *                 <-> nn.Module.load_state_dict() of those classes; keys and
 *                     shapes are exactly the reference state_dict layout
 *                     (trainer_seg_video_model.py:744-846 loads them)
 *   dmc_frame_stats
 *                 <-> caller-side metrics: rate_distortion_loss / _roi_mse /
 *                     _psnr_from_mse   trainer_seg_video_model.py:598-601,655-660,904-934
 *
 * Conventions
 *   - plain C, no exceptions: every int-returning function returns 0 on success
 *     or a negative DMC_E_* code; dmc_last_error() gives the message.
 *   - every tensor pointer is a DEVICE pointer to contiguous fp32 in the
 *     reference's own layout (NCHW); the caller (torch) owns all of them.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Work is enqueued asynchronously; nothing here synchronises the host.
 *   - one engine handle per (variant, batch, height, width); a handle is
 *     thread-compatible (use it from one thread / one stream at a time).  It
 *     belongs to the CUDA device that was current in dmc_create; every later
 *     call switches to that device for its duration.
 *   - height and width must be multiples of 16 for `old`, `fast`, `mask_prop`
 *     and the intra model (y is replicate-padded to a multiple of 4 for the
 *     hyper path, models/common_model.py:68-72) and multiples of 64 for
 *     `performance`, which never pads (seg_video_model.py:331) -- dmc_create
 *     refuses other sizes.
 */
#ifndef DMC_B200_H_
#define DMC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct dmc_engine dmc_engine;

/* trainer_seg_video_model.py:478-495 (`dmc_variant`) + the intra model */
enum {
  DMC_VARIANT_OLD = 0,
  DMC_VARIANT_PERFORMANCE = 1,
  DMC_VARIANT_FAST = 2,
  DMC_VARIANT_MASK_PROP = 3,
  DMC_VARIANT_INTRA = 4
};

enum {
  DMC_OK = 0,
  DMC_E_INVALID = -1,   /* bad argument / unknown key / shape mismatch      */
  DMC_E_CUDA = -2,      /* a CUDA runtime or driver call failed             */
  DMC_E_STATE = -3,     /* weights missing or not finalised                 */
  DMC_E_UNSUPPORTED = -4
};

/* dmc_create flags */
enum {
  DMC_FLAG_SIMT_GEMM = 1,   /* run every contraction on the fp32 CUDA-core kernel
                               (validation backend) instead of tcgen05          */
  DMC_FLAG_KEEP_TAPS = 2,   /* keep intermediate tensors readable via dmc_get_tap */
  DMC_FLAG_RECON_SPLIT3 = 8  /* run recon_generation_net with the fp32-grade 3-term split product too.  By default
                                its contractions use plain fp16 operands (hi planes, 1 term, fp32 accumulate): x_hat of
                                a P frame never feeds a later symbol, and PSNR moves by < 1e-3 dB */
};

int dmc_create(int variant, int batch, int height, int width, int flags, dmc_engine** out);
void dmc_destroy(dmc_engine* e);
/* message of the last failure on this handle (or of the last failed dmc_create if e==NULL) */
const char* dmc_last_error(const dmc_engine* e);

/* The state_dict keys this engine consumes (a subset of the reference's: unused
 * tensors such as hyper_in_adapter.* are not listed). */
int dmc_num_weights(const dmc_engine* e);
const char* dmc_weight_key(const dmc_engine* e, int index);
/* shape4 receives up to 4 dims; returns ndim or a negative error */
int dmc_weight_shape(const dmc_engine* e, int index, int64_t* shape4);

/* Hand one state_dict tensor (fp32, device, contiguous) to the engine.  Unknown keys
 * return DMC_E_INVALID; the data is repacked on `stream` into the engine's own
 * split-fp16 tiles (hi + 2^11-scaled lo), the caller's buffer is not retained. */
int dmc_set_weight(dmc_engine* e, const char* key, const float* dev_ptr,
                   const int64_t* shape, int ndim, void* stream);
int dmc_finalize_weights(dmc_engine* e, void* stream);

/* One P-frame.  x: (B,3,H,W).  mask: (B,1,H,W) or NULL (3-channel call).
 * dpb_frame: (B,3,H,W), read iff after_i != 0.  dpb_feature: (B,256,H/8,W/8), read iff
 * after_i == 0.  Outputs: x_hat (B,3,H,W) in [0,1]; feature (B,256,H/8,W/8);
 * bpp3 = B x {bpp, bpp_y, bpp_z}; mask_pred (B,1,H,W) or NULL -- written only by
 * mask_prop with after_i == 0 (the predictor's logits); finite_flag (int32, device)
 * or NULL: the reference's _finite_check sites (seg_video_model_fast.py:152-156,269-276,353-371) without their
 * host syncs -- one fused check at the end of the frame.  The word is zeroed and bit i is set if tensor i holds a
 * non-finite value or hit the fp16 range limit of the split storage format (|x| >= 65504 saturates):
 *   0 temporal feature (feature_adaptor_i / _p)   1 feature_extractor.ctx   2 feature_extractor.ctx_t
 *   3 y (encoder, after FiLM for `performance`)   4 z (hyper_encoder)       5 params (y_prior_fusion)
 *   6 y_hat (compress_prior_2x)                   7 decoder feature (dpb["feature"]) */
#define DMC_FINITE_TAGS "feature_adaptor", "feature_extractor.ctx", "feature_extractor.ctx_t", "encoder", \
                        "hyper_encoder", "y_prior_fusion", "y_hat", "decoder" 
int dmc_forward(dmc_engine* e, const float* x, const float* mask, const float* dpb_frame,
                const float* dpb_feature, int qp, int after_i, float* x_hat, float* feature,
                float* bpp3, float* mask_pred, int32_t* finite_flag, void* stream);

/* One I-frame through DMCI (engine created with DMC_VARIANT_INTRA). */
int dmci_forward(dmc_engine* e, const float* x, int qp, float* x_hat, float* bpp3, void* stream);

/* Decoder side (the split of DMC.forward / DMCI.forward that the reference sketches in its dead compress / decompress,
 * src/models/video_model.py:256-333, image_model.py): a decoder knows the dpb, qp and the decoded symbols, not x.
 *   dmc_decode_begin    temporal context (P engines) + hyper decoder + prior fusion from z_hat (B,128,H/64,W/64)
 *   for step in 0 .. steps-1   (2 checkerboard steps for P frames, 4 for the intra model):
 *     dmc_decode_sigma    runs the spatial prior of the step (step > 0) and writes the predicted scale of every element
 *                         this step owns into sigma_out (B,C,H/16,W/16; other elements are left stale)
 *     dmc_decode_symbols  takes the step's decoded symbols (same dense shape, read at the owned elements only)
 *   dmc_decode_finish   y_hat -> decoder -> feature (P engines) and x_hat
 * The same kernels and buffers as forward(): x_hat / feature are bit-identical to the encoder's. */
int dmc_decode_begin(dmc_engine* e, const float* dpb_frame, const float* dpb_feature, int qp, int after_i,
                     const float* z_hat, void* stream);
int dmc_decode_sigma(dmc_engine* e, int step, float* sigma_out, void* stream);
int dmc_decode_symbols(dmc_engine* e, int step, const float* symbols, void* stream);
int dmc_decode_finish(dmc_engine* e, float* x_hat, float* feature, void* stream);

/* Intermediate tensors of the last forward, converted to NCHW fp32 (needs
 * DMC_FLAG_KEEP_TAPS).  shape4 receives (B,C,H,W); dst may be NULL to query the shape. */
int dmc_get_tap(dmc_engine* e, const char* name, float* dst, int64_t capacity_elems,
                int64_t* shape4, void* stream);

/* Fused caller-side statistics.  Adds to stats7 (7 doubles, device):
 *   [0] sum bits_y  [1] sum bits_z  [2] sum (x_hat-x)^2  [3] sum m*(x_hat-x)^2
 *   [4] sum m (x3 channels)  [5] number of elements B*3*H*W  [6] number of frames (B)
 * m = (mask > 0).  mask may be NULL (ROI terms untouched). bpp3 as produced by dmc_forward. */
int dmc_frame_stats(double* stats7, const float* x_hat, const float* x, const float* mask,
                    const float* bpp3, int batch, int height, int width, void* stream);

/* Device data path (the dataset side of the caller): decoded camera frames, uint8 interleaved (frames, height, width, 3)
 * in R,G,B order (bgr = 1: B,G,R as cv2.imdecode returns them), and cached masks uint8 (frames, height, width) or
 * NULL  ->  (frames, out_channels, crop_h, crop_w) fp32 planes [Y, Cb, Cr(, mask)], cropped at (top, left).
 *   src/dataset/seg_waymo_dataset.py:26-34  _rgb_from_proto        uint8 / 255.0
 *   src/dataset/seg_waymo_dataset.py:36-43  _rgb_to_ycbcr_bt709    BT.709, clamp to [0, 1]
 *   src/dataset/seg_waymo_dataset.py:56-79  _load_cached_mask      mask = value > mask_threshold (0 for the 0/1 npz
 *                                                                  cache, 127 for the png cache); NULL mask = zeros
 *   src/dataset/seg_waymo_dataset.py:231-245 __getitem__           one crop for all frames, mask as channel 4
 * Bit-identical to those lines run on the CPU in fp32.  All pointers are device pointers. */
int dmc_frames_from_u8(const uint8_t* img, const uint8_t* mask, float* out, int frames, int height, int width, int top,
                       int left, int crop_h, int crop_w, int out_channels, int bgr, int mask_threshold, void* stream);

/* Mask propagation (BASELINE config 4): mask[i] = logits[i] > 0 ? 1 : 0 -- the MaskPredictor's logits of frame t-1
 * (dmc_forward's mask_pred, src/refactor/mask_predictor.py:27-46) become the mask frame t is coded with.
 * Both pointers 16-byte aligned. */
int dmc_mask_from_logits(const float* logits, float* mask, int64_t n, void* stream);

/* ---- single-operator entry points (the same kernels the engine launches), used by the
 * per-layer parity tests.  All tensors NCHW fp32 on the device. ---- */

/* act: 0 none, 1 WSiLU (layers.py:8-10), 2 ReLU.  backend: 0 tcgen05 split-fp16, 1 SIMT fp32.
 * nsplit: 1 = single term (fp16 hi planes only), any other value (3 by convention) = fp32-grade 3-term split
 * product.  groups must be 1 or cin (depthwise 3x3). */
int dmc_op_conv2d(const float* x, const float* weight, const float* bias, float* out, int batch,
                  int cin, int height, int width, int cout, int ksize, int stride, int padding,
                  int groups, int act, int nsplit, int backend, void* stream);
/* DepthConvBlock (layers.py:43-79); weights in the order adaptor(w,b or NULL,NULL), dc.0, dc.2,
 * dc.3, ffn.0, ffn.2; quant_step: (cout) or NULL. */
int dmc_op_depth_conv_block(const float* x, const float* const* weights12, const float* quant_step,
                            float* out, int batch, int cin, int cout, int height, int width,
                            int shortcut, int nsplit, int backend, void* stream);
/* formula 0: models/common_model.py:36-42, 1: refactor/common_model.py:37-68 */
int dmc_op_gaussian_bits(const float* sym, const float* sigma, float* bits, int64_t n, int formula,
                         void* stream);

/* ---- training mode (STE / noise quantisation, autograd through the fused DepthConvBlock) ----
 * What trainer_seg_video_model.py:983-1206 needs from the blocks when p_frame_model.train() is on.  A handle owns the
 * buffers, packed weights and launch programs of ONE DepthConvBlock geometry (layers.py:43-79); forward is the frame
 * engine's block, backward recomputes the block's intermediates from x (nothing else is kept between the passes) and
 * returns the gradients torch.autograd would: of x, of the twelve parameters (same order and shapes as weights12;
 * NULL entries are skipped) and of quant_step.  terms: 3 = fp32-grade split product, 1 = plain fp16 operands.
 * The gradient arithmetic runs on fp16 split planes; backward scales grad_out by a power of two (max |g| -> 2^8,
 * found on the device) on the way in and un-scales every result, so gradients of any magnitude are fine.  `out` is the
 * tensor forward returned (needed for grad_quant_step only, else it may be NULL).  All tensors NCHW fp32 on the
 * device, 16-byte aligned; calls are asynchronous on `stream`. */
typedef struct dmc_dcb_train dmc_dcb_train;
int dmc_dcb_train_create(int batch, int height, int width, int cin, int cout, int force_adaptor, int shortcut,
                         int has_quant_step, int terms, dmc_dcb_train** out);
void dmc_dcb_train_destroy(dmc_dcb_train* t);
const char* dmc_dcb_train_last_error(const dmc_dcb_train* t);
/* weights_unchanged != 0: the caller vouches that the twelve parameters hold the VALUES this handle packed at its
 * previous call (forward followed by backward of one training step): only packed copies still missing are made. */
int dmc_dcb_train_forward(dmc_dcb_train* t, const float* x, const float* const* weights12, const float* quant_step,
                          float* out, int weights_unchanged, void* stream);
int dmc_dcb_train_backward(dmc_dcb_train* t, const float* x, const float* const* weights12, const float* quant_step,
                           const float* out, const float* grad_out, float* grad_x, float* const* grad_weights12,
                           float* grad_quant_step, int weights_unchanged, void* stream);
/* A plain 1x1 convolution (stride 1, no padding, no groups) with its backward pass: the nn.Conv2d(cin, cout, 1) layers of
 * the models outside the DepthConvBlocks (feature_adaptor_p, encoder.conv1, decoder.proj, the recon head, sub-pixel and
 * prior heads).  Same conventions as the block entries above; weight (cout, cin, 1, 1), bias (cout) or NULL; any of the
 * gradient destinations may be NULL (x is only needed when grad_weight is wanted). */
typedef struct dmc_conv1x1_train dmc_conv1x1_train;
int dmc_conv1x1_train_create(int batch, int height, int width, int cin, int cout, int has_bias, int terms,
                             dmc_conv1x1_train** out);
void dmc_conv1x1_train_destroy(dmc_conv1x1_train* t);
const char* dmc_conv1x1_train_last_error(const dmc_conv1x1_train* t);
int dmc_conv1x1_train_forward(dmc_conv1x1_train* t, const float* x, const float* weight, const float* bias, float* out,
                              int weights_unchanged, void* stream);
int dmc_conv1x1_train_backward(dmc_conv1x1_train* t, const float* x, const float* weight, const float* grad_out,
                               float* grad_x, float* grad_weight, float* grad_bias, int weights_unchanged, void* stream);
/* k x k convolutions (kernel 2 or 3, stride 1 or 2, padding 0 or 1, no groups: encoder.down, mask_sft.down, the sub-pixel
 * 3x3 of decoder.up, the 2x2 stride-2 downs) with their backward pass, through the im2col view.  Conventions as above;
 * weight (cout, cin, k, k). */
typedef struct dmc_convkxk_train dmc_convkxk_train;
int dmc_convkxk_train_create(int batch, int height, int width, int cin, int cout, int ksize, int stride, int padding,
                             int has_bias, int terms, dmc_convkxk_train** out);
void dmc_convkxk_train_destroy(dmc_convkxk_train* t);
const char* dmc_convkxk_train_last_error(const dmc_convkxk_train* t);
int dmc_convkxk_train_forward(dmc_convkxk_train* t, const float* x, const float* weight, const float* bias, float* out,
                              int weights_unchanged, void* stream);
int dmc_convkxk_train_backward(dmc_convkxk_train* t, const float* x, const float* weight, const float* grad_out,
                               float* grad_x, float* grad_weight, float* grad_bias, int weights_unchanged, void* stream);
/* AdaptiveQuant in training mode (layers/inference.py:16-27).  mode 0 "ste": out = round(x) (the straight-through
 * gradient is the identity); mode 1 "noise": out = x + noise with noise ~ U(-half_bin, half_bin) drawn by the caller. */
int dmc_op_quant_train(const float* x, const float* noise, float* out, int64_t n, int mode, void* stream);
/* Gradient of dmc_op_gaussian_bits with respect to the symbols and sigma (same formula numbers). */
int dmc_op_gaussian_bits_backward(const float* sym, const float* sigma, const float* grad_bits, float* grad_sym,
                                  float* grad_sigma, int64_t n, int formula, void* stream);

/* ---- measurement support (bench.py) ---- */
/* number of kernels this library has launched in this process so far */
int64_t dmc_kernel_launches(void);
/* While enabled, every contraction launch of this engine is bracketed by CUDA events on its
 * stream.  dmc_profile_read synchronises, then returns the summed device time of those launches
 * (ms), their count, and their algorithmic FLOPs (2*M*N*K with the convolution's logical sizes),
 * and clears the record. */
int dmc_profile_enable(dmc_engine* e, int on);
int dmc_profile_read(dmc_engine* e, double* gemm_ms, int64_t* gemm_launches, double* gemm_flops,
                     double* issued_flops);

/* Times one contraction launch in isolation (bench / profiling tool): rows x k -> n on zero-filled
 * S3 buffers, `iters` launches between two CUDA events.  mode: 0 plain, 1 +WSiLU, 2 +residual,
 * 3 chunk-add pair layout (n = 4C, writes n/2 columns).  pair: 0 = general kernel on one CTA, 1 = its
 * cta_group::2 variant, 2 = the specialised CTA-pair kernel of gemm_s3.cu.
 * probe: 0 normal; bit0 no TMA loads, bit1 no MMA issue, bit2 no epilogue global traffic, bit3 no
 * epilogue at all (gemm_s3 only). */
int dmc_bench_gemm(int rows, int k, int n, int mode, int nsplit, int pair, int iters, int probe,
                   float* ms_per_launch);

/* Times the depthwise 3x3 kernel alone on zero-filled buffers (bench / profiling tool); f32_in selects the
 * variant that reads fp32 rows (the one DepthConvBlock uses) instead of S3 planes. */
int dmc_bench_dwconv(int batch, int height, int width, int channels, int f32_in, int iters,
                     float* ms_per_launch);

/* Times a stack of `blocks` DepthConvBlocks (zero weights and inputs) the way a frame runs them: chained
 * contraction launches + depthwise kernels.  Reports milliseconds per block (bench / profiling tool). */
int dmc_bench_dcb(int batch, int height, int width, int cin, int cout, int blocks, int iters,
                  float* ms_per_block);

int dmc_num_sms(void);
const char* dmc_version(void);

/* Accumulate-truncation compensation of the tcgen05 contractions (csrc/kernels.cu: acc_comp_scaled): kappa in units of
 * 2^-24 per MMA accumulate step; 0 switches it off.  Default 0.276 (measured), or the environment variable
 * DMC_ACC_COMP.  Engines created after the call use the new value (diagnostics / A-B runs). */
int dmc_set_acc_comp(float kappa);
float dmc_get_acc_comp(void);

/* ---- range coder (rANS) for the quantised latents: the arithmetic-coding half of the reference's entropy models,
 *   EntropyCoder / RansEncoder / RansDecoder     src/models/entropy_models.py:11-81
 *   GaussianEncoder.encode_y / decode_y          src/models/entropy_models.py:227-341
 *   BitEstimator.encode_z / decode_z             src/models/entropy_models.py:152-224
 * The reference's native coder (MLCodec_extensions_cpp) is not in its tree; this one restates the published algorithm
 * (byte-wise rANS, 16-bit cdfs, escape + 4-bit bypass groups) with many independent streams of 256 symbols so that a
 * frame codes in parallel.  Container: u32 n | u32 streams | u16 bytes[streams] | streams... (csrc/rans.cu).
 * Tables are built on the host exactly as the reference's Python does (GaussianEncoder.update, BitEstimator.update)
 * and handed over as int32 cdfs: cdf[n_cdf][stride], cdf_len[n_cdf] (= pmf length + 2), offset[n_cdf]. ---- */
typedef struct dmc_rans dmc_rans;
int dmc_rans_create(const int32_t* cdf_host, const int32_t* cdf_len_host, const int32_t* offset_host, int n_cdf,
                    int stride, dmc_rans** out);
void dmc_rans_destroy(dmc_rans* r);
const char* dmc_rans_last_error(const dmc_rans* r);
/* cdf index per symbol, device arrays: y from the predicted scale (build_index_enc / _dec, src/layers/inference.py:63-84:
 * clamp to [scale_min, scale_max], nearest entry of the log-spaced table); z from the channel of a flattened NCHW
 * tensor (BitEstimator.build_indexes: base = qp * channels). */
int dmc_rans_index_gaussian(const float* sigma, int64_t n, float scale_min, float scale_max, int levels, int32_t* idx,
                            void* stream);
int dmc_rans_index_channels(int64_t n, int64_t per_channel, int channels, int base, int32_t* idx, void* stream);
/* upper bound of the container size for n symbols */
int64_t dmc_rans_max_bytes(int64_t n);
/* symbols: device fp32, integer valued (what AdaptiveQuant produces).  Both calls synchronise the stream. */
int dmc_rans_encode(dmc_rans* r, const float* sym, const int32_t* idx, int64_t n, uint8_t* out, int64_t cap,
                    int64_t* nbytes, void* stream);
int dmc_rans_decode(dmc_rans* r, const uint8_t* in, int64_t nbytes, const int32_t* idx, int64_t n, float* sym_out,
                    void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DMC_B200_H_ */
